"""BASELINE.json configs[0], [2], [3] at their full frame sizes on ONE GPU (the multi-GPU ones as one rank's share plus the
property that makes the shards composable).  Prints a markdown table.  configs[1] is bench.py, configs[4] scripts/sweep_k1.py;
parity of these paths against the reference is the test suite's job (tests/test_gpu_parity.py: clip_720p is configs[0]).

    python scripts/run_configs.py            (on the GPU box)
"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import avd_b200.synth as synth
from avd_b200 import dist as avd_dist, fusion, hints as avd_hints, scoring, video
from avd_b200.engine import Engine

rows = []


def close(a, b, atol=5e-5):
    return abs(a - b) <= 1e-4 * abs(b) + atol


def timed(fn, reps=3):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    return (time.perf_counter() - t0) / reps, out


eng = Engine(0)

# ---- configs[0]: single 720p 30 fps 10 s clip (300 frames -> 20 samples), whole analyzer tail, vs the CPU body ----------
clip = synth.Clip(300, 720, 1280, seed=0, fps=30.0)
frames = clip.sampled(15)
pin = eng.pinned((len(frames), 720, 1280, 3)); pin[:] = frames
batch = [pin[i] for i in range(len(frames))]
meta = {"fps": 30.0, "width": 1280, "height": 720, "duration": 10.0}
dt, res = timed(lambda: video.analyze_frames(batch, meta, engine=eng))
res2 = video.analyze_frames(list(frames), meta, engine=eng)            # pageable copies of the same frames
ok = res == res2 and len(res["timeline"]) == 10
rows.append(("[0] 720p 30fps 10s clip, 20 samples, full tail", f"{1e3*dt:.2f} ms = {len(frames)/dt:.0f} frames/s",
             "-", "deterministic, timeline length = round(duration): %s" % ok))
assert ok

# ---- configs[2]: 4K 60 fps 30 s clip (1800 frames -> 60 samples); rank 3 of 8's share with its halo vs the whole clip ----------
clip = synth.Clip(1800, 2160, 3840, seed=2, fps=60.0)
n = 60
pin = eng.pinned((n, 2160, 3840, 3))
for i in range(n):
    pin[i] = clip[i * 30]
allf = [pin[i] for i in range(n)]
dt_all, full = timed(lambda: eng.analyze_host(allf), reps=2)
lo, hi, halo = avd_dist.local_range_with_halo(n, 3, 8)
dt_sh, part = timed(lambda: eng.analyze_host([pin[halo]] + allf[lo:hi]), reps=3)
same = part[1:].tobytes() == full[lo:hi].tobytes()
rows.append(("[2] 4K 60fps 30s clip, 60 samples (1.49 GB)", f"1 GPU, whole clip: {1e3*dt_all:.1f} ms = {n/dt_all:.0f} frames/s ({n*2160*3840*3/dt_all/1e9:.1f} GB/s host->device)",
             f"rank 3 of 8 (samples {lo}..{hi-1} + halo): {1e3*dt_sh:.1f} ms", "shard records byte-identical to the whole-clip run: %s" % same))
assert same
del pin, allf

# ---- configs[3]: 256 clips x 720p x 5 s (10 samples each); one rank's share at 8 GPUs = 32 clips ----------------------------------
ids = list(avd_dist.shard_clips(256, 5, 8))
pin = eng.pinned((len(ids) * 10, 720, 1280, 3))
clips, metas = [], []
for j, cid in enumerate(ids):
    c = synth.Clip(150, 720, 1280, seed=1000 + cid, fps=30.0)
    for s_i in range(10):
        pin[j * 10 + s_i] = c[s_i * 15]
    clips.append([pin[j * 10 + s_i] for s_i in range(10)])
    metas.append({"fps": 30.0, "width": 1280, "height": 720, "duration": 5.0})
dt_b, outs = timed(lambda: video.analyze_frames_batch(clips, metas, engine=eng), reps=3)
one = video.analyze_frames(clips[7], metas[7], engine=eng)
same = outs[7] == one
rows.append(("[3] 256 x 720p 5s clips, clip-sharded x8: one rank's 32 clips (320 samples)", f"{1e3*dt_b:.1f} ms = {320/dt_b:.0f} frames/s per GPU",
             "-", "batched result of a clip == that clip analysed alone: %s" % same))
assert same

# the same 32 clips through to fused verdicts: host tail (records -> NumPy tail -> fuse per clip) against K3 (records stay in HBM)
hn = [avd_hints.compute_hints(m) for m in metas]
def host_tail():
    vs = video.analyze_frames_batch(clips, metas, engine=eng)
    return [{"video": v, "fusion": fusion.fuse(fusion.neutral_audio(5), v, h)} for v, h in zip(vs, hn)]
def device_tail():
    return scoring.score_clip_batch(eng, clips, metas, audios=[fusion.neutral_audio(5) for _ in clips], hints=hn)
dt_h, out_h = timed(host_tail, reps=3)
dt_d, out_d = timed(device_tail, reps=3)
same = out_h == out_d
rows.append(("[3] same 32 clips, frames -> fused verdicts", f"host tail {1e3*dt_h:.1f} ms = {320/dt_h:.0f} frames/s; device tail (K3) {1e3*dt_d:.1f} ms = {320/dt_d:.0f} frames/s",
             "-", "device tail == host tail (every float equal): %s" % same))
assert same

print("| config | avd-b200 (host frames -> results, one B200) | one rank's share | check |")
print("|---|---|---|---|")
for r in rows:
    print("| " + " | ".join(r) + " |")
eng.close()

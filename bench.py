#!/usr/bin/env python
"""bench.py -- analyzed frames/sec of the per-frame analyzer hot path at 1080p (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--configs]

One step = one pass of the hot path over one batch of synthetic input:
  N=1   BASELINE.json configs[1]: a 1080p 30 fps 60 s synthetic clip (1800 frames -> 120 sampled frames, 746 MB of
        BGR, larger than the 126 MB L2 so nothing is cache-resident between steps): K1 (per-frame features) + K2
        (119 Farneback pairs) + record read-back + the scalar tail (timeline, summary, fuse).
  N>1   weak scaling: an N x 60 s clip, frame-chunk sharded -- every rank analyses 120 sampled frames plus a
        one-sample halo, then one NCCL all-gather of the 160-byte records; rank 0 runs the tail.
The clip exists as the decoder's pictures (dense planar 4:2:0, what FFmpeg hands OpenCV) and as the BGR frames
cv2.VideoCapture.retrieve() makes of them (video.py:32; libswscale's arithmetic): both arms analyse those BGR frames.
`value`   BGR frames resident in HBM when the timed region starts (device entry points of the C ABI, CUDA-graph replay);
          records are read back and the tail runs on the host every step, one step behind the device (pipelined).
`e2e`     the same metric through the public host call (Engine.analyze_host): the decoder's 4:2:0 pictures in pinned
          host memory, host->device copies, the on-device colour conversion and the record read-back inside the timed
          region (`e2e.bgr_upload` is the same with BGR host frames, 2x the bytes, as in round 1).
`roofline` K1 (the HBM-bound kernel north_star puts the >=90 % target on): algorithmic bytes = H*W*3 per frame
           x frames per launch / mean launch duration from CUDA events on the launching stream, against
          MEASURED_PEAKS.json.  `roofline_k2`: the Farneback kernels (the rest of the step) against their own DRAM traffic.
`parity`  verdicts computed in this run: GPU records vs the reference's own analyze() golden (tests/golden/
          golden_1080p_full.json, same clip) and vs the reference's cv2/NumPy body run on this box's CPU on the same frames.
`cpu_baseline` / --impl reference: the reference's per-sample body on the box's host cores (oracle/cv_body.py:
          the same cv2/NumPy calls as video.py:36-57; /root/reference itself is not on the GPU box).
"""
import argparse
import json
import os
import sys
import threading
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, W, FPS, CLIP_FRAMES, SEED = 1080, 1920, 30.0, 1800, 1
STEP = 15                              # video.py:19 at 30 fps
SAMPLES = CLIP_FRAMES // STEP          # 120 analysed frames per 60 s clip
FRAME_BYTES = H * W * 3
YUV_BYTES = H * W * 3 // 2
METRIC = "analyzed frames/sec at 1080p"
WORKLOAD = ("1080p 30fps 60s synthetic clip (1800 frames -> 120 sampled frames) per GPU: "
            "per-frame features + 119 Farneback pairs + timeline/fuse")        # the same string in both arms
MIN_TIMED_S = 0.5                      # the K timed steps are repeated in rounds until the timed region is at least this long


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def make_inputs(rank: int, world: int, yuv_out: np.ndarray, raw_out=None, first_k=None):
    """This rank's 60 s chunk of the (world x 60 s) seed-1 clip as decoder pictures: yuv_out[0] = halo sample (ranks > 0),
    yuv_out[1 + k] = sample k.  raw_out (optional): the synthetic BGR frames themselves (golden check)."""
    import avd_b200.synth as synth
    clip = synth.Clip(CLIP_FRAMES * world, H, W, SEED, fps=FPS)
    clip._texture(0); clip._texture(1)                        # build both textures once (shared by the worker threads)
    first = rank * SAMPLES
    n = SAMPLES if first_k is None else first_k

    def one(k):
        f = clip[(first + k) * STEP]
        if raw_out is not None:
            raw_out[k] = f
        yuv_out[1 + k] = synth.bgr_to_i420(f)

    with ThreadPoolExecutor(min(8, host_cores())) as ex:
        list(ex.map(one, range(n)))
    if rank > 0:
        yuv_out[0] = synth.bgr_to_i420(clip[(first - 1) * STEP])


def decoded_bgr_cpu(yuv: np.ndarray, out: np.ndarray):
    """What cv2.VideoCapture.retrieve() returns for the pictures (oracle of libswscale's conversion) -- CPU arm input."""
    from oracle import yuv_oracle

    def one(k):
        out[k] = yuv_oracle.i420_to_bgr(yuv[k], H, W)

    with ThreadPoolExecutor(min(16, host_cores())) as ex:
        list(ex.map(one, range(len(yuv))))


def profile_json(name):
    p = os.path.join(ROOT, "profiles", name)
    return json.load(open(p)) if os.path.exists(p) else None


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """SM clock / throttle reasons during the timed region (NVML; nvidia-smi as a fallback)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples, self.reasons, self.max_mhz = index, False, [], set(), None
        self.nv = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        names = {0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown",
                 0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown", 0x10: "sync_boost"}
        while not self.stop_flag:
            try:
                if self.nv:
                    self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                    r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(self.nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                        else self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                    for bit, nm in names.items():
                        if r & bit:
                            self.reasons.add(nm)
                else:
                    import subprocess
                    o = subprocess.run(["nvidia-smi", f"--id={self.index}", "--query-gpu=clocks.sm,clocks.max.sm,clocks_event_reasons.active",
                                        "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout.split(",")
                    self.samples.append(int(o[0])); self.max_mhz = int(o[1])
            except Exception:
                pass
            time.sleep(0.004)

    def result(self):
        self.stop_flag = True
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s)}


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU path (cv2/NumPy port of video.py:36-57) on all host cores, same frames."""
    if rank != 0:
        return
    from oracle import cv_body
    cores = host_cores()
    yuv = np.empty((SAMPLES + 1, YUV_BYTES), np.uint8)
    make_inputs(0, 1, yuv)
    frames = np.empty((SAMPLES, H, W, 3), np.uint8)
    decoded_bgr_cpu(yuv[1:], frames)
    del yuv
    pool = cv_body.Pool(frames, cores)
    for _ in range(args.warmup):
        pool.step()
    t0 = time.perf_counter()
    done = 0
    for _ in range(args.steps):
        done += pool.step()
    dt = time.perf_counter() - t0
    pool.close()
    v = done / dt
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8/f32 (OpenCV)", "data": "synthetic",
            "config": {"workload": WORKLOAD, "frames_per_step": SAMPLES, "runs_on": "host CPU"},
            "cpu_baseline": {"value": v, "unit": "frames/s", "cores": pool.procs, "kind": "port",
                             "sample": f"{args.steps} x the full 120-sample clip, one process per core, cv2.setNumThreads(1), decode excluded"},
            "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---- sub-results for BASELINE configs[2] and configs[3] (N > 1, or --configs) -------------------------------------------------
def sub_config2(eng, rank, world, dev, barrier, reduce_max, reduce_min):
    """configs[2]: 4K 60 fps 30 s clip (1800 frames -> 60 samples), frame-chunk sharded with a one-sample halo and an
    all-gather of the records (avd_b200.dist.analyze_frames_sharded, host frames in pinned memory)."""
    import avd_b200.synth as synth
    from avd_b200 import dist as avd_dist
    from oracle import cv_body, parity
    H4, W4, n_total, step = 2160, 3840, 60, 30
    clip = synth.Clip(1800, H4, W4, 2, fps=60.0)
    lo, hi, halo_i = avd_dist.local_range_with_halo(n_total, rank, world)
    pin = eng.pinned((hi - lo + 1, H4, W4, 3))
    for k, i in enumerate(range(lo, hi)):
        pin[1 + k] = clip[i * step]
    if halo_i is not None:
        pin[0] = clip[halo_i * step]
    local = [pin[1 + k] for k in range(hi - lo)]
    halo = pin[0] if halo_i is not None else None
    meta = {"fps": 60.0, "width": W4, "height": H4, "duration": 30.0}
    run = lambda: avd_dist.analyze_frames_sharded(local, halo, n_total, meta, rank, world, eng.analyze_host, device=dev, H=H4, W=W4)
    for _ in range(2):
        res, full = run()
    iters = 4
    barrier(); t0 = time.perf_counter()
    for _ in range(iters):
        res, full = run()
    barrier(); dt = reduce_max(time.perf_counter() - t0)
    # parity: (i) this rank's first two samples (seeded by its halo) against the reference's body on the CPU, integer
    # fields included; (ii) the gathered chain is consistent across every shard boundary
    m = min(2, hi - lo)
    cpu = cv_body.run_chunk(local[:m], halo, collect_ints=True)
    keys = ("textures", "flow_means", "flow_vars", "dup", "timeline", "hash_words", "hams", "lap_sums")
    mine = parity.records_vs_cpu_body(full[lo:lo + m], dict(zip(keys, cpu)), H4 * W4, first_has_prev=halo is not None)
    # (iii) every sample of this rank's shard, and on rank 0 the returned dict, against the reference's own analyze() run on the
    # whole clip (tests/golden/golden_4k_full.json)
    gp = os.path.join(ROOT, "tests", "golden", "golden_4k_full.json")
    gold, vs_gold, dict_ok = (json.load(open(gp)) if os.path.exists(gp) else None), None, None
    if gold is not None:
        vs_gold = parity.records_vs_golden_range(full[lo:hi], gold, lo, first_has_prev=halo is not None)
        dict_ok = parity.video_dict_vs_golden(res, gold["result"])
    ok = reduce_min(1.0 if (mine["ok"] and parity.chain_consistent(full) and (vs_gold is None or (vs_gold["ok"] and dict_ok))) else 0.0) > 0.5
    eng.free_pinned(pin)
    return {"workload": "configs[2]: 4K 60fps 30s clip, 60 sampled frames (1.49 GB BGR) frame-chunk sharded, 1-sample halo, all-gather of records",
            "value": n_total * iters / dt, "unit": "frames/s", "ms_per_pass": 1e3 * dt / iters, "samples_per_rank": hi - lo,
            "parity": {"ok": bool(ok), "rank0_first2_vs_cpu_body": mine, "rank0_shard_vs_reference_golden": vs_gold, "result_dict_vs_reference_golden": dict_ok,
                       "checked": "every rank: all samples of its shard vs the reference's analyze() golden of the whole clip (ints exact, flow in tolerance) and the "
                                  "assembled result dict vs the reference's; first 2 samples after the shard boundary vs the CPU body; gathered Hamming chain consistent"},
            "ai_score": None if res is None else float(np.mean(res["timeline"]))}


def sub_config3(eng, rank, world, barrier, reduce_max, reduce_min, clips_per_rank=32):
    """configs[3]: 256 short 720p clips (5 s -> 10 samples each) clip-sharded: 32 clips per GPU, frames -> fused verdicts
    (scoring.score_clip_batch: K1 + K2 + the device scoring tail K3), host frames in pinned memory."""
    import avd_b200.synth as synth
    from avd_b200 import fusion, hints, scoring, video
    from oracle import cv_body, parity
    Hc, Wc = 720, 1280
    seeds = [1000 + rank * clips_per_rank + i for i in range(clips_per_rank)]
    pin = eng.pinned((clips_per_rank, 10, Hc, Wc, 3))

    def gen(i):
        pin[i] = synth.Clip(150, Hc, Wc, seeds[i], fps=30.0).sampled(15)

    with ThreadPoolExecutor(min(8, max(1, host_cores() // max(1, world)))) as ex:
        list(ex.map(gen, range(clips_per_rank)))
    clips = [[pin[i, j] for j in range(10)] for i in range(clips_per_rank)]
    meta = {"fps": 30.0, "width": Wc, "height": Hc, "duration": 5.0, "bit_rate": 2_500_000}
    metas = [dict(meta) for _ in clips]
    hn = [hints.compute_hints(dict(meta), "") for _ in clips]
    run = lambda: scoring.score_clip_batch(eng, clips, metas, audios=[fusion.neutral_audio(5) for _ in clips], hints=hn, with_raw=True)
    for _ in range(2):
        out = run()
    iters = 4
    barrier(); t0 = time.perf_counter()
    for _ in range(iters):
        out = run()
    barrier(); dt = reduce_max(time.perf_counter() - t0)
    ok, worst = True, 0.0
    for ci in (0, clips_per_rank - 1):                         # two clips per rank against the reference's body on the CPU
        tex, fm, fv, dup, tl = cv_body.run_chunk(clips[ci])
        v = out[ci]["video"]
        t = parity.timeline_vs_cpu(v["timeline"][:5], tl[:5])
        ref_v = {"timeline": tl[:5], "timeline_ai": tl[:5],
                 "summary": {"dup_density": float(dup / 9), "scene_change_rate": float(np.mean(np.array(fv) > 0.5)), "flow_mean": float(np.mean(fm)),
                             "flow_var": float(np.var(fm)), "texture_var": float(np.var(tex)), "w": Wc, "h": Hc, "fps": 30.0}}
        ref_f = fusion.fuse(fusion.neutral_audio(5), ref_v, hn[ci], with_raw=True)
        d = abs(ref_f["_raw_score"] - out[ci]["fusion"]["_raw_score"])
        worst = max(worst, d)
        ok &= (t["ok"] and d <= 1e-3 and ref_f["result"]["label"] == out[ci]["fusion"]["result"]["label"]
               and v["summary"]["dup_density"] == ref_v["summary"]["dup_density"]
               and abs(v["summary"]["texture_var"] - ref_v["summary"]["texture_var"]) <= 1e-9 * abs(ref_v["summary"]["texture_var"]))
    # every clip of this rank against the reference's analyze() + fuse() run on the same clip (tests/golden/golden_clips256.json)
    gp = os.path.join(ROOT, "tests", "golden", "golden_clips256.json")
    n_gold = 0
    if os.path.exists(gp):
        gold = {c["seed"]: c for c in json.load(open(gp))["clips"]}
        for ci, sd in enumerate(seeds):
            if sd in gold:
                n_gold += 1
                ok &= parity.video_dict_vs_golden(out[ci]["video"], gold[sd]["video"]) and parity.fused_vs_golden(out[ci]["fusion"], gold[sd]["fused"])
    ok = reduce_min(1.0 if ok else 0.0) > 0.5
    eng.free_pinned(pin)
    return {"workload": f"configs[3]: {clips_per_rank * world} x 720p 5s clips (10 samples each) clip-sharded, {clips_per_rank} clips per GPU, frames -> fused verdicts (K1+K2+K3)",
            "value": clips_per_rank * 10 * world * iters / dt, "unit": "frames/s", "ms_per_pass": 1e3 * dt / iters,
            "clips_per_s": clips_per_rank * world * iters / dt,
            "parity": {"ok": bool(ok), "score_delta_max_rank0": worst, "clips_vs_reference_golden_rank0": n_gold,
                       "checked": "every clip vs the reference's analyze()+fuse() golden (timeline, summary, label, peaks, binned timeline, rounded scores); "
                                  "2 clips per rank vs the CPU body + host fuse: timeline in tolerance, dup_density equal, texture_var 1e-9, pre-rounding score <= 1e-3, label equal"}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--configs", action="store_true", help="also run the configs[2] / configs[3] sub-results at N=1 (always on for N>1)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank, world)
    if args.warmup < 3 and not os.environ.get("AVD_BENCH_ALLOW_SHORT"):
        args.warmup = 3                                     # timing rules: at least 3 warm-up steps (the line reports what was run)
    args.steps = max(1, args.steps)

    numa_bound = False
    if world > 1 and os.environ.get("AVD_BIND_NUMA", "1") != "0":
        from avd_b200 import dist as avd_dist     # staging memory on the GPU's own NUMA node (first touch after binding)
        numa_bound = avd_dist.bind_to_gpu_numa(local_rank)
    # ---- synthetic input (host, before CUDA is touched so the CPU baseline may fork) ---------------------
    host_yuv = np.empty((SAMPLES + 1, YUV_BYTES), np.uint8)
    raw = np.empty((SAMPLES, H, W, 3), np.uint8) if (rank == 0 and world == 1) else None
    make_inputs(rank, world, host_yuv, raw)

    from oracle import cv_body, parity as par
    cpu_baseline, cpu_out, cpu_ints = None, None, None
    if rank == 0:
        n_dec = SAMPLES if (world == 1 and not args.no_cpu_baseline) else 16
        dec = np.empty((n_dec, H, W, 3), np.uint8)
        decoded_bgr_cpu(host_yuv[1:1 + n_dec], dec)
        if world == 1 and not args.no_cpu_baseline:
            cores = host_cores()
            pool = cv_body.Pool(dec, cores)
            pool.step()                                        # warm (page-in, cv2 init)
            t0 = time.perf_counter(); n_done = pool.step(); dt_all = time.perf_counter() - t0
            cpu_out = pool.outputs()                           # the timed pass's own outputs: the parity check reads them
            pool.close()
            t0 = time.perf_counter(); cv_body.run_chunk(dec[:16]); dt_one = time.perf_counter() - t0
            cpu_baseline = {"value": n_done / dt_all, "unit": "frames/s", "cores": pool.procs, "kind": "port",
                            "sample": "the full 120-sample clip once (after one warm pass), one process per core, cv2.setNumThreads(1), decode excluded",
                            "single_process_value": 16 / dt_one}
        # integer intermediates of the reference's body for the first 16 samples (untimed)
        keys = ("textures", "flow_means", "flow_vars", "dup", "timeline", "hash_words", "hams", "lap_sums")
        cpu_ints = dict(zip(keys, cv_body.run_chunk(dec[:16], collect_ints=True)))
        dec_head = dec[:4].copy()
        del dec

    # ---- GPU --------------------------------------------------------------------------------------------
    import torch
    import torch.distributed as dist
    from avd_b200 import fusion, hints, video
    from avd_b200._lib import REC_DTYPE
    from avd_b200.engine import Engine
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    eng = Engine(local_rank)
    n_total = SAMPLES * world
    meta = {"fps": FPS, "width": W, "height": H, "duration": CLIP_FRAMES * world / FPS, "bit_rate": 8_000_000}
    hnt = hints.compute_hints(meta, "")

    pinned = eng.pinned((SAMPLES + 1, YUV_BYTES))          # slot 0 = halo
    pinned[...] = host_yuv
    del host_yuv
    d_yuv = torch.from_numpy(pinned).to(dev)
    d_all = eng.yuv420_to_bgr(d_yuv, H, W).view(SAMPLES + 1, H, W, 3)     # the decoded BGR frames, resident in HBM
    d_frames = d_all[1:]
    d_halo = d_all[0] if rank > 0 else None
    del d_yuv
    pinned_bgr = eng.pinned((SAMPLES + 1, H, W, 3))        # the same frames as BGR host buffers (e2e.bgr_upload)
    torch.from_numpy(pinned_bgr).copy_(d_all)
    torch.cuda.synchronize()

    def tail(recs_np):
        res = video.result_from_records(recs_np, H, W, FPS, meta["duration"], w=W, h=H)
        return fusion.fuse(fusion.neutral_audio(len(res["timeline"])), res, hnt, with_raw=True), res

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_(x, op):
        if world == 1:
            return float(x)
        t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=op)
        return float(t.item())

    reduce_max = lambda x: reduce_(x, dist.ReduceOp.MAX)
    reduce_min = lambda x: reduce_(x, dist.ReduceOp.MIN)

    # ---- device-resident step: graph replay -> (all-gather) -> async read-back; host tail one step behind ----------------
    use_graph = os.environ.get("AVD_BENCH_GRAPH", "1") != "0"
    comm = torch.cuda.Stream()
    send = [torch.zeros((SAMPLES, REC_DTYPE.itemsize), dtype=torch.uint8, device=dev) for _ in range(2)]
    recv = [torch.empty((n_total, REC_DTYPE.itemsize), dtype=torch.uint8, device=dev) for _ in range(2)] if world > 1 else send
    host_rec = [torch.empty((n_total, REC_DTYPE.itemsize), dtype=torch.uint8).pin_memory() for _ in range(2)]
    ev_done = [torch.cuda.Event() for _ in range(2)]
    ev_host = [torch.cuda.Event() for _ in range(2)]

    def device_steps(steps, graph=None):
        out = None
        fn = eng.analyze_device_graph if (use_graph if graph is None else graph) else eng.analyze_device

        def finish(i):
            ev_host[i % 2].synchronize()
            return tail(host_rec[i % 2].numpy().view(REC_DTYPE).reshape(-1)) if rank == 0 else None

        for i in range(steps):
            b = i % 2
            send[b].copy_(fn(d_frames, prev_frame=d_halo))
            ev_done[b].record()
            with torch.cuda.stream(comm):
                comm.wait_event(ev_done[b])
                if world > 1:
                    dist.all_gather_into_tensor(recv[b], send[b])
                if rank == 0:
                    host_rec[b].copy_(recv[b], non_blocking=True)
                ev_host[b].record(comm)
            if i > 0:
                out = finish(i - 1)
        return finish(steps - 1) if steps else out

    from avd_b200.engine import HostBatch
    lo_slot = 0 if rank > 0 else 1                         # the staging buffers keep their addresses: pointers prepared once
    host_batch = HostBatch(pinned[lo_slot:], "i420", (H, W))
    host_batch_bgr = HostBatch(pinned_bgr[lo_slot:])
    send_pad = torch.zeros((SAMPLES, REC_DTYPE.itemsize), dtype=torch.uint8, device=dev)

    tail_pool = ThreadPoolExecutor(1)

    def host_steps(steps, bgr=False):
        # the O(samples) host tail of step i (NumPy / Python) runs on a worker thread while the main thread is inside the
        # library call of step i + 1 (which releases the GIL), as in the device-resident loop above; every step's tail
        # finishes inside the timed region (the last one is waited for before returning)
        pending = None
        for _ in range(steps):
            recs = eng.analyze_host(host_batch_bgr if bgr else host_batch)
            if rank > 0:
                recs = recs[1:]
            if world > 1:
                send_pad.copy_(torch.from_numpy(recs.view(np.uint8).reshape(SAMPLES, -1)))
                dist.all_gather_into_tensor(recv[0], send_pad)
                recs = recv[0].cpu().numpy().view(REC_DTYPE).reshape(-1) if rank == 0 else None
            if rank == 0:
                if pending is not None:
                    pending.result()
                pending = tail_pool.submit(tail, recs)
        return pending.result() if pending is not None else None

    def timed(fn, steps, rounds=1, prof=False):
        """`rounds` x `steps` steps between barrier + synchronize on both sides; device time from CUDA events on the
        launching stream, max with the wall clock, max over ranks."""
        barrier()
        if prof:
            eng.profile_read(); eng.profile(True)
        l0 = eng.launches + eng.graph_launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        out = None
        for _ in range(rounds):
            out = fn(steps)
        e1.record()
        barrier()
        wall = time.perf_counter() - t0
        ms = max(e0.elapsed_time(e1), 0.0)
        kern = None
        if prof:
            eng.profile(False); kern = eng.profile_read()
        return reduce_max(max(ms / 1e3, wall)), eng.launches + eng.graph_launches - l0, kern, out

    def rounds_for(fn):
        est, _, _, _ = timed(fn, args.steps)                  # untimed estimate (also extra warm-up)
        return max(1, int(np.ceil(MIN_TIMED_S / max(est, 1e-6))))

    device_steps(args.warmup)
    r_dev = rounds_for(device_steps)
    sampler = ClockSampler(local_rank); sampler.start()
    dt, launches, _, result = timed(device_steps, args.steps, r_dev)
    clocks_dev = sampler.result()
    # same loop again with per-launch CUDA events on (kept out of `value`: the events add ~1 us per launch)
    _, _, kern, _ = timed(lambda s: device_steps(s, graph=False), args.steps, prof=True)
    host_steps(args.warmup)
    r_e2e = rounds_for(host_steps)
    sampler = ClockSampler(local_rank); sampler.start()
    dt_e2e, _, _, result_e2e = timed(host_steps, args.steps, r_e2e)
    clocks_e2e = sampler.result()
    host_steps(2, bgr=True)
    dt_e2e_bgr, _, _, result_e2e_bgr = timed(lambda s: host_steps(s, bgr=True), args.steps)

    # two requests in flight per GPU, as the service's engine pool runs them (engine.lease): a second context uploads the next
    # request while the first one finishes its last K1/K2 -- the un-hideable tail of ONE request is hidden by the next one
    eng2 = Engine(local_rank)
    host_batch2 = HostBatch(pinned[lo_slot:], "i420", (H, W))

    def host_steps_2(steps):
        import concurrent.futures as cf
        with cf.ThreadPoolExecutor(2) as ex:
            futs = [ex.submit((eng if i % 2 == 0 else eng2).analyze_host, host_batch if i % 2 == 0 else host_batch2) for i in range(steps)]
            out = None
            for f in futs:
                recs = f.result()
                if rank > 0:
                    recs = recs[1:]
                out = tail(recs) if (rank == 0 and world == 1) else None
        return out

    host_steps_2(4)
    dt_e2e_2, _, _, result_e2e_2 = timed(host_steps_2, args.steps)
    eng2.close()

    # the machine's host->device ceiling for this many GPUs copying at once (plain cudaMemcpyAsync of the same pinned
    # buffers, no kernels): what `e2e` can at best reach on this host, since every analysed byte has to cross PCIe
    d_sink = torch.empty((SAMPLES + 1, YUV_BYTES), dtype=torch.uint8, device=dev)
    t_pin = torch.from_numpy(pinned)

    def h2d_only(steps):
        for _ in range(steps):
            d_sink.copy_(t_pin, non_blocking=True)

    h2d_only(2)
    dt_h2d, _, _, _ = timed(h2d_only, 8)
    h2d_ceiling_gbs = world * 8 * t_pin.numel() / dt_h2d / 1e9
    del d_sink

    # BASELINE configs[4] / SURVEY 8(d) quote the per-frame kernel at >= 256 frames per launch: the same kernel on 256 resident
    # frames, reported next to the 120-frame launch of the step (extra information, not the `roofline.frac` of the contract)
    k1_256 = None
    if world == 1:
        big = torch.cat([d_frames, d_frames, d_frames[:256 - 2 * SAMPLES]])
        for _ in range(3):
            eng.frames_features(big)
        torch.cuda.synchronize()
        eng.profile_read(); eng.profile(True)
        for _ in range(5):
            eng.frames_features(big)
        eng.profile(False)
        k = eng.profile_read().get("k1_fused")
        if k:
            ms = k["ms"] / k["launches"]
            k1_256 = {"frames_per_launch": 256, "ms_per_launch": ms, "achieved": FRAME_BYTES * 256 / (ms * 1e-3) / 1e9}
        del big

    # ---- parity verdicts of this run ---------------------------------------------------------------------------------------
    parity = {}
    e2e_recs = eng.analyze_host(host_batch)
    if rank > 0:
        e2e_recs = e2e_recs[1:]
    dev_recs = eng.records_to_numpy(eng.analyze_device(d_frames, prev_frame=d_halo))
    same_paths = e2e_recs.tobytes() == dev_recs.tobytes()
    if rank == 0:
        parity["k0_vs_swscale_oracle_exact"] = bool(np.array_equal(d_frames[:4].cpu().numpy(), dec_head))
        parity["e2e_records_equal_device_records"] = bool(same_paths)
        parity["first16_vs_cpu_body_ints"] = par.records_vs_cpu_body(e2e_recs[:16], cpu_ints, H * W)
        if cpu_out is not None:                                # N=1: the timed CPU pass covered all 120 samples
            parity["all120_vs_cpu_body"] = par.records_vs_cpu_body(e2e_recs, cpu_out, H * W)
            parity["timeline_vs_cpu_body"] = par.timeline_vs_cpu(result_e2e[1]["timeline"], cpu_out["timeline"])
        gp = os.path.join(ROOT, "tests", "golden", "golden_1080p_full.json")
        g = json.load(open(gp)) if os.path.exists(gp) else None
        if world == 1 and g is not None and raw is not None:
            d_raw = torch.from_numpy(raw).to(dev)
            raw_recs = eng.records_to_numpy(eng.analyze_device(d_raw))
            del d_raw
            parity["raw_clip_vs_reference_golden"] = par.records_vs_reference_golden(raw_recs, g)
            f_raw, v_raw = tail(raw_recs)
            ref_f = fusion.fuse(fusion.neutral_audio(len(g["result"]["timeline"])), json.loads(json.dumps(g["result"])), g["hints"], with_raw=True)
            parity["raw_clip_result_vs_reference_golden"] = par.result_vs_reference_golden(
                v_raw, f_raw["_raw_score"], g["result"], ref_f["_raw_score"], g["fused_with_neutral_audio"], f_raw)
    if world > 1:
        # rank 0 recomputes the head of rank 1's chunk on its own GPU (same frames, same halo) and compares with what rank 1
        # contributed to the all-gather: sharding changes no arithmetic
        gathered = None
        if rank == 0:
            gathered = recv[0].cpu().numpy().view(REC_DTYPE).reshape(-1).copy()      # last e2e step's all-gather
            k = 24
            y1 = np.empty((SAMPLES + 1, YUV_BYTES), np.uint8)
            make_inputs(1, world, y1, first_k=k)
            d1 = eng.yuv420_to_bgr(torch.from_numpy(y1[:k + 1]).to(dev), H, W).view(k + 1, H, W, 3)
            r1 = eng.records_to_numpy(eng.analyze_device(d1[1:], prev_frame=d1[0]))
            parity["rank1_head_recomputed_on_rank0_equal"] = bool(r1.tobytes() == gathered[SAMPLES:SAMPLES + k].tobytes())
            parity["gathered_chain_consistent"] = bool(par.chain_consistent(gathered))
            del d1, y1
    all_same = reduce_min(1.0 if same_paths else 0.0) > 0.5
    if rank == 0:
        parity["e2e_records_equal_device_records_all_ranks"] = bool(all_same)
        checks = [v if isinstance(v, bool) else v["ok"] for v in parity.values()]
        parity["int_exact"] = bool(all(v["int_exact"] for v in parity.values() if isinstance(v, dict) and "int_exact" in v))
        parity["max_rel_float"] = min(1e30, max([v.get("flow_max_rel", 0.0) for v in parity.values() if isinstance(v, dict)] + [0.0]))
        parity["score_delta"] = parity.get("raw_clip_result_vs_reference_golden", {}).get("score_delta")
        parity["ok"] = bool(all(checks))
        parity["gates"] = "ints bit-exact; floats |d| <= 1e-4|ref| + 1e-6 (5e-5 px for flow statistics of static pairs); pre-rounding score |d| <= 1e-3"

    # ---- sub-results: configs[2] and configs[3] ----------------------------------------------------------------------------
    subs = None
    if world > 1 or args.configs:
        del d_all, d_frames, pinned_bgr
        torch.cuda.empty_cache()
        subs = [sub_config2(eng, rank, world, dev, barrier, reduce_max, reduce_min),
                sub_config3(eng, rank, world, barrier, reduce_max, reduce_min)]

    clocks = dict(clocks_dev, samples_e2e=clocks_e2e["samples"], sm_mhz_e2e=clocks_e2e["sm_mhz"],
                  reasons=sorted(set(clocks_dev["reasons"]) | set(clocks_e2e["reasons"])))

    if rank == 0:
        timed_steps = args.steps * r_dev
        value = n_total * timed_steps / dt
        hbm, hbm_src = peaks()
        k1 = kern.get("k1_fused") or {"ms": sum(kern[k]["ms"] for k in ("gray", "lap", "area_h", "area_v", "area_int", "thumb") if k in kern),
                                      "launches": kern.get("gray", {}).get("launches", 1)}
        k1_name = "k1_fused" if "k1_fused" in kern else "k1_generic(gray+lap+area+thumb)"
        frames_per_launch = (SAMPLES + (1 if d_halo is not None else 0)) * args.steps / max(1, k1["launches"])
        k1_ms = k1["ms"] / max(1, k1["launches"])
        achieved = FRAME_BYTES * frames_per_launch / (k1_ms * 1e-3) / 1e9
        total_kernel_ms = sum(v["ms"] for v in kern.values())
        tr1 = profile_json("r02_k1_traffic.json") or profile_json("r01_k1_traffic.json")
        k1_traffic = (tr1["dram_bytes_read"] + tr1["dram_bytes_write"]) / tr1["frames_per_launch"] * frames_per_launch if tr1 and (tr1["H"], tr1["W"]) == (H, W) else None
        k2_names = [k for k in kern if k in ("pyr", "polyexp", "update_matrices", "box_solve", "flow_iter", "magstats")]
        k2_ms_events = sum(kern[k]["ms"] for k in k2_names) / args.steps       # launches on two streams overlap: an upper bound
        # K2's share of the timed step itself: the graph-replayed step minus the per-frame kernels (event-timed, serial)
        k2_ms = min(k2_ms_events, 1e3 * dt / (args.steps * r_dev) - sum(v["ms"] for k, v in kern.items() if k not in k2_names) / args.steps)
        n_pairs = SAMPLES - (0 if d_halo is not None else 1)
        tr2 = profile_json("r02_k2_traffic.json")
        roofline_k2 = None
        if tr2:
            traffic = float(tr2["dram_bytes_per_step"]) * n_pairs / tr2["pairs_per_step"]
            roofline_k2 = {"kernels": k2_names, "bound": "hbm (its own DRAM traffic; algorithmic input is 204 800 B per pair)", "ms_per_step": k2_ms, "ms_per_step_sum_of_launches": k2_ms_events, "us_per_pair": 1e3 * k2_ms / n_pairs,
                           "traffic": traffic, "traffic_source": "ncu --set full, dram__bytes_read+write summed over the step's Farneback launches (profiles/r02_k2_traffic.json)",
                           "achieved": traffic / (k2_ms * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s", "frac": traffic / (k2_ms * 1e-3) / 1e9 / hbm,
                           "algorithmic_bytes": 2 * 102400 * n_pairs, "share_of_step": k2_ms / (1e3 * dt / (args.steps * r_dev))}
        f_e2e, v_e2e = result_e2e
        line = {
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / timed_steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8 (K0/K1 integer path) / f32+f64 (Farneback)", "data": "synthetic",
            "config": {"workload": WORKLOAD, "frames_per_step_per_gpu": SAMPLES,
                       "sharding": "none" if world == 1 else f"frame-chunk x{world}, 1-sample halo, all-gather of 160 B records",
                       "l2_policy": "inputs (746 MB/GPU) larger than L2 (126 MB), no flush needed", "seed": SEED, "numa_bound": numa_bound,
                       "device_path": "CUDA graph replay of the step's kernel launches" if use_graph else "plain stream launches",
                       "timed_region": f"{r_dev} rounds x {args.steps} steps = {dt:.3f} s (value), {r_e2e} rounds x {args.steps} steps = {dt_e2e:.3f} s (e2e)"},
            "rounds": r_dev, "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": n_total * args.steps * r_e2e / dt_e2e, "unit": "frames/s",
                    "h2d_bytes_per_step": int(YUV_BYTES * (SAMPLES * world + world - 1)), "d2h_bytes_per_step": int(160 * n_total),
                    "ms_per_step": 1e3 * dt_e2e / (args.steps * r_e2e), "rounds": r_e2e,
                    "host_format": "planar 4:2:0 (the decoder's pictures, 1.5 B/px); BGR conversion of cv2.VideoCapture.retrieve() on the device",
                    "host_tail": "timeline / fuse of step i on a worker thread while step i + 1 is inside the library call; the last one is waited for inside the timed region",
                    "h2d_gbs_per_gpu": YUV_BYTES * (SAMPLES + (1 if world > 1 else 0)) / (dt_e2e / (args.steps * r_e2e)) / 1e9,
                    "h2d_ceiling_gbs": h2d_ceiling_gbs,
                    "h2d_ceiling_note": f"{world} GPU(s) copying the same pinned buffers concurrently, no kernels (aggregate GB/s); e2e moves "
                                        f"{YUV_BYTES * (SAMPLES * world + world - 1) / (dt_e2e / (args.steps * r_e2e)) / 1e9:.1f} GB/s",
                    "frac_of_h2d_ceiling": YUV_BYTES * (SAMPLES * world + world - 1) / (dt_e2e / (args.steps * r_e2e)) / 1e9 / h2d_ceiling_gbs,
                    "two_requests_in_flight": {"value": n_total * args.steps / dt_e2e_2, "ms_per_step": 1e3 * dt_e2e_2 / args.steps,
                                               "note": "two contexts per GPU alternate (the service's engine pool); per-rank records only, no all-gather"
                                                       + ("" if world > 1 else "; result equals the sequential one: " + str(result_e2e_2[0] == result_e2e[0]))},
                    "bgr_upload": {"value": n_total * args.steps / dt_e2e_bgr, "h2d_bytes_per_step": int(FRAME_BYTES * (SAMPLES * world + world - 1)),
                                   "ms_per_step": 1e3 * dt_e2e_bgr / args.steps}},
            "roofline": {"kernel": k1_name, "bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm,
                         "traffic": k1_traffic if k1_name == "k1_fused" else None,
                         "traffic_source": "ncu --set full, dram__bytes_read+write per launch (profiles/r0x_k1_traffic.json)", "peak_source": hbm_src,
                         "ms_per_launch": k1_ms, "frames_per_launch": frames_per_launch,
                         "share_of_kernel_time": k1["ms"] / total_kernel_ms if total_kernel_ms else None,
                         "at_256_frames_per_launch": dict(k1_256, frac=k1_256["achieved"] / hbm) if k1_256 else None},
            "roofline_k2": roofline_k2,
            "kernels": {k: {"ms_per_step": v["ms"] / args.steps, "launches_per_step": v["launches"] / args.steps} for k, v in sorted(kern.items())},
            "k2_us_per_pair": 1e3 * k2_ms / n_pairs,
            "cpu_baseline": cpu_baseline,
            "parity": parity,
            "sub_results": subs,
            "result": {"ai_score": f_e2e["result"]["ai_score"], "label": f_e2e["result"]["label"],
                       "e2e_matches_device": f_e2e == result[0], "bgr_upload_matches": result_e2e_bgr[0] == f_e2e},
        }
        print(json.dumps(line), flush=True)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""CPU: host-side mirror of the reference interface (video tail, hints, fusion), the C-ABI surface, and the
sharding logic over gloo (world_size 2).  No GPU compute."""
import copy
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from conftest import FLOW_ATOL, ROOT, close


def _records_from_golden(c):
    from avd_b200._lib import REC_DTYPE
    n = c["n_samples"]
    r = np.zeros(n, REC_DTYPE)
    r["lap_sum"], r["lap_sumsq"] = c["lap_sum"], c["lap_sumsq"]
    r["flow_mean"][1:] = [m for m, _ in c["flow_stats"]]
    r["flow_var"][1:] = [v for _, v in c["flow_stats"]]
    r["flags"][1:] = 3
    r["ham_prev"][0] = -1
    r["ham_prev"][1:] = c.get("ham", [7] * (n - 1))
    r["hash"] = np.array(c["hash_words"], np.uint32)
    return r


def _same_result(out, ref, tex_rtol=1e-12):
    assert out["timeline"] is out["timeline_ai"]                     # video.py:83 returns one list twice
    assert len(out["timeline"]) == len(ref["timeline"])
    assert all(isinstance(x, float) for x in out["timeline"])        # JSON-serialisable plain floats
    for a, b in zip(out["timeline"], ref["timeline"]):
        assert abs(a - b) <= 1e-12 + tex_rtol * abs(b)
    for k, b in ref["summary"].items():
        a = out["summary"][k]
        assert type(a) is type(b)
        if isinstance(b, int) or k in ("dup_density", "scene_change_rate", "fps"):
            assert a == b, k
        else:
            assert abs(a - b) <= 1e-10 * abs(b), (k, a, b)


@pytest.mark.parametrize("name", ["clip_small", "clip_odd", "clip_nometa", "clip_720p", "clip_1080p_head"])
def test_video_tail_matches_reference(golden, name):
    from avd_b200 import video
    c = golden["clips"][name]
    dur = c["meta"].get("duration") or c["n_frames"] / c["fps"]
    out = video.result_from_records(_records_from_golden(c), c["H"], c["W"], c["fps"], dur)
    _same_result(out, c["result"])


def test_video_tail_edge_cases():
    from avd_b200 import video
    from avd_b200._lib import REC_DTYPE
    r = video.result_from_records(np.zeros(0, REC_DTYPE), 0, 0, 0.0, 0.0)        # nothing decoded
    assert r["timeline"] == [0.5] and r["summary"]["dup_density"] == 0.0 and r["summary"]["flow_mean"] == 0.0
    one = np.zeros(1, REC_DTYPE); one["lap_sumsq"] = 100 * 64 * 64; one["ham_prev"] = -1
    r = video.result_from_records(one, 64, 64, 30.0, 3.4)                        # one sample, padded to round(3.4)=3
    assert len(r["timeline"]) == 3 and r["timeline"][0] == r["timeline"][2] == 1.0 - 100.0 / 1100.0
    assert video.sample_step(25) == 12 and video.sample_step(None) == 15 and video.sample_step(1) == 1
    assert video.result_from_records(np.zeros(0, REC_DTYPE), 0, 0, 0.0, 2.5)["timeline"] == [0.5, 0.5]   # round half even


def test_analyze_unopenable_path_returns_reference_shape():
    from avd_b200 import video
    assert video.analyze("/nonexistent/clip.mp4", {}) == {"timeline": [], "summary": {}, "timeline_ai": []}   # video.py:12-13


def test_fusion_and_hints_match_reference(golden):
    from avd_b200 import fusion, hints
    for c in golden["fuse"]:
        audio, video = copy.deepcopy(c["audio"]), copy.deepcopy(c["video"])
        assert hints.compute_hints(c["meta"], "x") == c["hints"]
        assert fusion.fuse(audio, video, c["hints"]) == c["out"]
        assert audio["timeline"] == c["audio_timeline_after"] and video["timeline"] == c["video_timeline_after"]


def test_fusion_score_rounding_gate(golden):
    from avd_b200 import fusion
    c = golden["fuse"][3]
    out = fusion.fuse(copy.deepcopy(c["audio"]), copy.deepcopy(c["video"]), c["hints"], with_raw=True)
    assert abs(out["_raw_score"] - c["out"]["result"]["ai_score"]) <= 0.005 + 1e-12


def test_c_abi_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "avd.h")).read()
    declared = set(re.findall(r"\b(avd_[a-z0-9_]+)\s*\(", hdr))
    from avd_b200 import _lib
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    so = _lib.LIB_PATH
    if not os.path.exists(so):
        import __graft_entry__ as g
        g.build()
    L = C.CDLL(so)
    for s in declared:
        assert hasattr(L, s), s
    assert b"sm_100a" in _lib.load().avd_version()
    # the SASS in the library is sm_100a only
    out = subprocess.run(["cuobjdump", "-lelf", so], capture_output=True, text=True).stdout
    assert "sm_100a" in out and "sm_90" not in out


def test_plain_c_client_builds_and_fails_loudly_without_gpu():
    """tests/native/abi_smoke.c: the C ABI is usable from plain C (gcc, no C++/CUDA/torch headers).  Without a GPU the
    client must report AVD_ERR_NODEVICE and exit 2 -- no CPU fallback behind the boundary."""
    import subprocess
    csrc = os.path.join(ROOT, "avd_b200", "csrc")
    subprocess.check_call(["make", "-s", "-C", csrc, "abi_smoke"])
    exe = os.path.join(ROOT, "tests", "native", "abi_smoke")
    assert os.path.exists(exe)
    import ctypes
    lib = ctypes.CDLL(os.path.join(csrc, "libavd.so"))
    if lib.avd_device_count() <= 0:
        out = subprocess.run([exe], capture_output=True, text=True)
        assert out.returncode == 2 and "no CPU fallback" in out.stderr


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from avd_b200._lib import AvdError
    from avd_b200.engine import Engine
    with pytest.raises(AvdError):
        Engine(0)
    from avd_b200 import video
    with pytest.raises(AvdError):
        video.analyze_frames([np.zeros((64, 64, 3), np.uint8)], {"fps": 30.0})


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "avd_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle|liboracle|#include.*oracle|dlopen", src, re.M), \
                    f"{f} reaches into oracle/"


def test_shard_bounds():
    from avd_b200 import dist
    assert [hi - lo for lo, hi in dist.shard_bounds(60, 8)] == [8, 8, 8, 8, 7, 7, 7, 7]
    assert dist.shard_bounds(3, 4) == [(0, 1), (1, 2), (2, 3), (3, 3)]
    assert dist.local_range_with_halo(120, 0, 2) == (0, 60, None)
    assert dist.local_range_with_halo(120, 1, 2) == (60, 120, 59)
    assert dist.local_range_with_halo(3, 3, 4) == (3, 3, None)
    assert list(dist.shard_clips(256, 7, 8)) == list(range(224, 256))


_WORKER = r"""
import os, sys, json, numpy as np
sys.path.insert(0, {root!r})
import torch.distributed as dist
from avd_b200 import dist as avd_dist, video
from avd_b200._lib import REC_DTYPE
import avd_b200.synth as synth
from oracle import oracle as O

def records_fn(frames):       # CPU stand-in for Engine.analyze_host: same record layout, oracle arithmetic
    out = np.zeros(len(frames), REC_DTYPE)
    for i, r in enumerate(O.frames_features(frames)):
        out[i]["lap_sum"], out[i]["lap_sumsq"] = r["lap_sum"], r["lap_sumsq"]
        out[i]["hash"] = r["hash"]; out[i]["ham_prev"] = r["ham"]; out[i]["flags"] = 1 if r["ham"] >= 0 else 0
        if "flow_mean" in r:
            out[i]["flow_mean"], out[i]["flow_var"] = r["flow_mean"], r["flow_var"]; out[i]["flags"] |= 2
    return out

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
clip = synth.Clip(105, 96, 128, seed=31, fps=30.0)
frames = clip.sampled(15)                      # 7 samples -> shards 4 + 3
meta = {{"fps": 30.0, "width": 128, "height": 96, "duration": 3.5}}
lo, hi, halo = avd_dist.local_range_with_halo(len(frames), rank, world)
res, full = avd_dist.analyze_frames_sharded(frames[lo:hi], frames[halo] if halo is not None else None,
                                            len(frames), meta, rank, world, records_fn)
single = records_fn(list(frames))
assert full.tobytes() == single.tobytes(), "sharded records differ from the single-process run"
want = video.result_from_records(single, 96, 128, 30.0, 3.5)
assert res == want
if rank == 0:
    print("SHARD_OK", json.dumps(res["summary"]))
dist.destroy_process_group()
"""


def test_frame_chunk_sharding_world2_gloo(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_WORKER.format(root=ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29531", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "SHARD_OK" in outs[0]


# ---- per-shape plan of the frame kernels (host code of libavd.so, no device needed) --------------------------------
class _Plan(C.Structure):
    _fields_ = [("area_int", C.c_int), ("thumb_2x2", C.c_int), ("strip_ntx_max", C.c_int), ("k1_path", C.c_int),
                ("row_owner_begin", C.c_int * 33), ("thumb_row_begin", C.c_int * 33),
                ("lx_i0", C.c_int * 320), ("lx_a0", C.c_int * 320), ("lx_a1", C.c_int * 320),
                ("ly_i0", C.c_int * 320), ("ly_a0", C.c_int * 320), ("ly_a1", C.c_int * 320),
                ("strip_tx0", C.c_int * 256), ("strip_ntx", C.c_int * 256), ("rowinfo", C.c_uint32 * 4096)]


def _plan(H, W):
    from avd_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    L = C.CDLL(_lib.LIB_PATH)
    L.avd_shape_plan.argtypes = [C.c_int, C.c_int, C.POINTER(_Plan)]
    p = _Plan()
    assert L.avd_shape_plan(H, W, C.byref(p)) == 0
    return p


def _lin_taps(ssize, dsize=320):
    """INTER_LINEAR 8-bit taps as the library builds them (float32 arithmetic), x-axis form."""
    scale = 1.0 / (dsize / ssize)
    out = []
    for d in range(dsize):
        f = np.float32((d + 0.5) * scale - 0.5)
        s = int(np.floor(f))
        f = np.float32(f - np.float32(s))
        if s < 0:
            f, s = np.float32(0), 0
        if s >= ssize - 1:
            f, s = np.float32(0), ssize - 1
        out.append((s, int(np.rint(np.float32(np.float32(1) - f) * np.float32(2048))), int(np.rint(f * np.float32(2048)))))
    return out


@pytest.mark.parametrize("shape", [(720, 1280), (1080, 1920), (1440, 2560), (2160, 3840), (480, 854), (480, 640), (648, 1344)])
def test_shape_plan_taps_bands_and_strips(shape):
    H, W = shape
    p = _plan(H, W)
    assert [(p.lx_i0[i], p.lx_a0[i], p.lx_a1[i]) for i in range(320)] == _lin_taps(W)
    ly = _lin_taps(H)
    assert [(p.ly_i0[i], p.ly_a0[i], p.ly_a1[i]) for i in range(320)] == ly      # downscaling: no row clipping in play
    # bands: every source row has exactly one owner, band b starts at ceil(b*H/32)
    own = list(p.row_owner_begin)
    assert own[0] == 0 and own[32] == H and all(own[b] == -(-b * H // 32) for b in range(32)) and own == sorted(own)
    # every thumbnail row belongs to the band that owns its top tap row
    trb = list(p.thumb_row_begin)
    assert trb[0] == 0 and trb[32] == 320
    for b in range(32):
        for ty in range(trb[b], trb[b + 1]):
            assert own[b] <= p.ly_i0[ty] < own[b + 1]
    # every thumbnail column belongs to exactly one 16-pixel strip: the one that holds its left tap
    seen = []
    for s in range((W + 15) // 16):
        for k in range(p.strip_ntx[s]):
            tx = p.strip_tx0[s] + k
            assert 16 * s <= p.lx_i0[tx] < 16 * s + 16
            seen.append(tx)
    assert seen == list(range(320))
    assert p.strip_ntx_max == max(p.strip_ntx[s] for s in range((W + 15) // 16))


@pytest.mark.parametrize("shape", [(720, 1280), (1080, 1920), (2160, 3840), (648, 1344)])
def test_shape_plan_row_roles(shape):
    """Thumbnail role of every source row (streaming kernel): top / bottom tap of exactly one thumbnail row."""
    H, W = shape
    p = _plan(H, W)
    want = [0] * H
    for ty in range(320):
        y0 = p.ly_i0[ty]
        want[y0] = 1 | (p.ly_a0[ty] << 2) | (ty << 16)
        want[y0 + 1] = 3 | (p.ly_a1[ty] << 2) | (ty << 16)
        assert p.ly_a0[ty] + p.ly_a1[ty] == 2048
    assert list(p.rowinfo[:H]) == want


def test_shape_plan_kernel_choice():
    """Which per-frame kernel a dense aligned buffer gets: 0 streaming (TMA) kernel, 3 any-width kernel (a band of gray rows in
    shared memory; round 2: faster than the banded kernel on every shape both take), 2 banded kernel (W % 16 == 0 bands too
    large for the any-width kernel's shared memory), 1 generic (ragged widths with such bands)."""
    want = {(1080, 1920): 0, (720, 1280): 0, (1440, 2560): 0, (2160, 3840): 0, (648, 1344): 0, (960, 1280): 0,
            (480, 640): 3, (640, 640): 3, (360, 640): 3, (1080, 1936): 3,   # W % 16 == 0 but below / off the streaming kernel's grid
            (480, 854): 3, (333, 777): 3, (768, 1366): 3, (32, 32): 3, (33, 35): 3,
            (600, 1920): 3,                                                   # H < 640: a tap row would have two roles
            (2160, 3856): 2, (2160, 3838): 1}
    for (H, W), k in want.items():
        assert _plan(H, W).k1_path == k, (H, W)
    assert _plan(960, 1280).area_int == 1 and _plan(1080, 1920).area_int == 0 and _plan(640, 640).thumb_2x2 == 1


def test_scoring_job_layout_matches_header(tmp_path):
    """The NumPy dtypes scoring.py packs jobs / reads scores with are the C structs of include/avd.h."""
    from avd_b200 import scoring
    src = tmp_path / "lay.c"
    fields_j = [n for n in scoring.JOB_DTYPE.names]
    fields_s = [n for n in scoring.SCORE_DTYPE.names]
    body = "".join(f'printf("j %s %zu\\n", "{f}", offsetof(avd_clip_job, {f}));' for f in fields_j)
    body += "".join(f'printf("s %s %zu\\n", "{f}", offsetof(avd_clip_score, {f}));' for f in fields_s)
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "avd.h"\nint main(){' + body +
                   'printf("J %zu\\nS %zu\\n", sizeof(avd_clip_job), sizeof(avd_clip_score));return 0;}')
    exe = tmp_path / "lay"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split("\n")
    for line in out:
        t = line.split()
        if not t:
            continue
        if t[0] == "j":
            assert scoring.JOB_DTYPE.fields[t[1]][1] == int(t[2]), line
        elif t[0] == "s":
            assert scoring.SCORE_DTYPE.fields[t[1]][1] == int(t[2]), line
        elif t[0] == "J":
            assert scoring.JOB_DTYPE.itemsize == int(t[1])
        elif t[0] == "S":
            assert scoring.SCORE_DTYPE.itemsize == int(t[1])


def test_scoring_joblist_offsets_and_timeline_length():
    from avd_b200 import scoring
    assert [scoring.timeline_length(d) for d in (0.0, 0.4, 0.5, 1.5, 2.5, 3.5, 59.6)] == [1, 1, 1, 2, 2, 4, 60]   # half to even
    jl = scoring.JobList()
    jl.add(rec_first=0, rec_count=10, npix=64, tlen=5)                                                    # video only: 5 outputs
    jl.add(rec_first=10, rec_count=3, npix=64, tlen=2, audio={"timeline": [0.1] * 7, "flags_audio": {}}, hints={"bpp": 0.5})
    jl.add(video={"timeline": [], "summary": {}}, audio={"timeline": []}, hints={})                       # both empty: L = 1
    jobs, tl_in, n_out = jl.arrays()
    assert jobs["out_off"].tolist() == [0, 5, 12] and n_out == 13 and len(tl_in) == 7
    assert jobs["do_fuse"].tolist() == [0, 1, 1] and jobs["video_given"].tolist() == [0, 0, 1]
    assert float(jobs["pen"][1]) == 0.0 and float(jobs["pen"][2]) == 0.05                                 # bpp default 0.0 < 0.07


def test_audio_tail_matches_reference(synth):
    """audio.result_from_features (the host tail of the audio drop-in) on the oracle's per-window features == the
    reference's dicts; and the failure result has the reference's shape (audio.py:112-118)."""
    import json
    from avd_b200 import audio
    from oracle import audio_oracle
    with open(os.path.join(ROOT, "tests", "golden", "golden_audio.json")) as f:
        G = json.load(f)
    for c in G["cases"]:
        wav = synth.audio(c["seed"], c["n"], c["kind"])
        with np.errstate(all="ignore"):
            feats = np.stack(audio_oracle.window_features(wav, c["sr"]), axis=1)
            assert audio.result_from_features(feats, len(wav), c["sr"]) == c["out"], c["name"]
    bad = audio.analyze("/nonexistent/clip.mp4", {"duration": 2.6})
    assert bad["scores"] == {} and "error" in bad["flags_audio"] and bad["timeline"] == [0.5, 0.5, 0.5]


class _OracleEngine:
    """CPU stand-in for Engine.analyze_host (same record layout, oracle arithmetic) for host-logic tests."""
    calls = 0

    def analyze_host(self, frames):
        from avd_b200._lib import REC_DTYPE
        from oracle import oracle as O
        type(self).calls += 1
        out = np.zeros(len(frames), REC_DTYPE)
        for i, r in enumerate(O.frames_features(list(frames))):
            out[i]["lap_sum"], out[i]["lap_sumsq"] = r["lap_sum"], r["lap_sumsq"]
            out[i]["hash"] = r["hash"]; out[i]["ham_prev"] = r["ham"]; out[i]["flags"] = 1 if r["ham"] >= 0 else 0
            if "flow_mean" in r:
                out[i]["flow_mean"], out[i]["flow_var"] = r["flow_mean"], r["flow_var"]; out[i]["flags"] |= 2
        return out


def test_request_batcher_coalesces_and_keeps_results(synth):
    """service.RequestBatcher: requests released together end up in fewer engine runs, each caller gets exactly the
    result it would have got alone; a request of another frame size goes its own way; an engine failure reaches every
    caller of the batch."""
    import threading
    from avd_b200 import service, video
    eng = _OracleEngine()
    clips = [synth.Clip(40, 64, 96, seed=50 + i, events=False) for i in range(4)]
    reqs = [([c[0], c[20]], {"fps": 30.0, "width": 96, "height": 64, "duration": 1.0 + i}) for i, c in enumerate(clips)]
    other = synth.Clip(40, 48, 64, seed=60, events=False)
    reqs.append(([other[0], other[20]], {"fps": 30.0, "duration": 2.0}))
    want = [video.analyze_frames(f, m, engine=eng) for f, m in reqs]
    _OracleEngine.calls = 0
    b = service.RequestBatcher(engine=eng, max_wait_ms=300.0)
    got, gate = [None] * len(reqs), threading.Barrier(len(reqs))

    def work(i):
        gate.wait()
        got[i] = b.analyze_frames(*reqs[i])

    ts = [threading.Thread(target=work, args=(i,)) for i in range(len(reqs))]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert got == want
    assert sorted(b.batches) == [1, 4] and _OracleEngine.calls == 2      # four same-size requests in one run, the odd one alone

    class Boom:
        def analyze_host(self, frames):
            raise RuntimeError("device lost")
    bad = service.RequestBatcher(engine=Boom(), max_wait_ms=50.0)
    errs = []

    def work2(i):
        try:
            bad.analyze_frames(*reqs[i])
        except RuntimeError as e:
            errs.append(str(e))
    ts = [threading.Thread(target=work2, args=(i,)) for i in range(3)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert errs == ["device lost"] * 3


def test_service_analyze_path_neutral_results_without_decoders(tmp_path):
    """service.analyze_path on a path nothing can open: the video side returns the reference's unopenable-file result
    (video.py:12-13), the audio side its failure result (audio.py:112-118); fuse still yields a verdict."""
    from avd_b200 import service
    out = service.analyze_path(str(tmp_path / "missing.mp4"), {"duration": 3.2, "fps": 30.0, "width": 640, "height": 360, "bit_rate": 500000})
    assert out["video"] == {"timeline": [], "summary": {}, "timeline_ai": []}      # fuse pads a copy of an empty list (fusion.py:18)
    assert "error" in out["audio"]["flags_audio"] and out["audio"]["timeline"] == [0.5, 0.5, 0.5]
    assert set(out) == {"hints", "video", "audio", "result", "timeline_binned", "peaks"}
    assert out["result"]["label"] in ("real", "uncertain", "ai") and len(out["timeline_binned"]) == 3


class _FakeEngine:
    """Stands in for Engine in host-logic tests: `analyze_host` 'analyses' a frame by reading its first byte; a record's
    ham_prev is the previous frame's first byte (or -1), so chains, halos and ordering are observable without a GPU."""

    def __init__(self):
        import threading
        self._lock = threading.RLock()
        self.calls, self.live = [], 0

    def staging(self, nbytes):
        import contextlib

        @contextlib.contextmanager
        def cm():
            self.live += 1
            try:
                yield np.zeros(int(nbytes), np.uint8)
            finally:
                self.live -= 1
        return cm()

    def analyze_host(self, frames, keep_on_device=False, fmt="bgr", shape=None):
        from avd_b200._lib import REC_DTYPE
        n = len(frames)
        first = [int(np.asarray(frames[i]).reshape(-1)[0]) for i in range(n)]
        self.calls.append((fmt, n, first))
        recs = np.zeros(n, REC_DTYPE)
        for i in range(n):
            recs["lap_sum"][i] = first[i]
            recs["ham_prev"][i] = first[i - 1] if i else -1
            recs["flags"][i] = 3 if i else 0
        return recs


def test_decode_run_pipeline_halo_order_and_staging():
    """video._RunPipeline (decode of run k+1 overlapped with the device work of run k): every sampled frame is analysed
    exactly once and in order, every run after the first is seeded with the previous run's last frame in its halo slot and
    that record is dropped, and the staging buffers are returned only after the worker is done."""
    from avd_b200 import video
    eng = _FakeEngine()
    nb, n = 6 * 4 * 3, 2 * video.DECODE_RUN + 7                 # tiny 6x4 BGR frames, two full runs and a partial one
    import contextlib
    with contextlib.ExitStack() as stack:
        stages = [stack.enter_context(eng.staging((video.DECODE_RUN + 1) * nb)) for _ in range(2)]
        pipe = video._RunPipeline(eng, stages, nb, "bgr", (6, 4))
        try:
            for i in range(n):
                pipe.slot()[:] = (i + 1) % 251
                pipe.pushed()
            recs = pipe.finish()
        finally:
            pipe.close()
        assert eng.live == 2
    assert eng.live == 0
    assert [int(x) for x in recs["lap_sum"]] == [(i + 1) % 251 for i in range(n)]
    assert int(recs["ham_prev"][0]) == -1
    assert [int(x) for x in recs["ham_prev"][1:]] == [(i + 1) % 251 for i in range(n - 1)]      # the chain crosses the runs
    assert [c[1] for c in eng.calls] == [video.DECODE_RUN, video.DECODE_RUN + 1, 8]               # runs 2 and 3 carry a halo frame
    assert eng.calls[1][2][0] == video.DECODE_RUN % 251 and eng.calls[2][2][0] == (2 * video.DECODE_RUN) % 251


def test_engine_pool_leases_and_bounds(monkeypatch):
    """engine.EnginePool: at most `size` contexts, a lease blocks until one is free, an explicit engine bypasses the pool."""
    import threading
    import time
    from avd_b200 import engine as E
    made = []

    class Fake:
        def __init__(self, device):
            made.append(self)

        def close(self):
            pass

    monkeypatch.setattr(E, "Engine", Fake)
    pool = E.EnginePool(0, size=2)
    got, gate = [], threading.Event()

    def work():
        with pool.lease() as e:
            got.append(e)
            gate.wait(2.0)

    ts = [threading.Thread(target=work) for _ in range(3)]
    [t.start() for t in ts]
    time.sleep(0.2)
    assert len(made) == 2 and len(got) == 2                      # the third caller waits for a context
    gate.set()
    [t.join() for t in ts]
    assert len(made) == 2 and len(got) == 3 and got[2] in made
    sentinel = object()
    with E.lease(sentinel) as e:
        assert e is sentinel
    pool.close()

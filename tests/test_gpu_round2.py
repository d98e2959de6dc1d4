"""GPU tests added in round 2 (pytest -m gpu): the decode -> device seam (4:2:0 upload + on-device colour conversion),
the full-size configs[1] golden, and the engine's resource handling (staging reuse, graph invalidation, engine pool).
Everything goes through the C ABI; the checkers are the CPU oracles and the fixtures made from the reference's own code."""
import copy
import hashlib
import json
import os
import threading

import numpy as np
import pytest

from conftest import ROOT, close, flow_atol, summary_flow_atol

pytestmark = pytest.mark.gpu


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def torch():
    import torch as t
    assert t.cuda.is_available(), "these tests need a GPU"
    return t


@pytest.fixture(scope="module")
def eng(torch):
    from avd_b200.engine import Engine
    e = Engine(0)
    yield e
    assert e.check_guards() == 0, "a kernel wrote outside one of the library's device workspaces"
    e.close()


def _i420(rng, H, W):
    return rng.integers(0, 256, H * W * 3 // 2, dtype=np.uint8)


@pytest.mark.parametrize("H,W", [(1080, 1920), (720, 1280), (480, 854), (34, 66), (2160, 3840)])
@pytest.mark.parametrize("full", [False, True])
def test_k0_colour_conversion_equals_swscale_oracle(torch, eng, H, W, full):
    """csrc/k0_yuv.cu against the restatement of libswscale's yuv420p -> bgr24 (itself pinned to the library): random
    planes, vector kernel (W % 16 == 0) and byte kernel (854, 66), limited and full range, dense and padded BGR pitch."""
    from oracle import yuv_oracle as YO
    rng = np.random.default_rng([H, W, int(full)])
    n = 2
    yuv = np.stack([_i420(rng, H, W) for _ in range(n)])
    want = np.stack([YO.i420_to_bgr(f, H, W, full) for f in yuv])
    for pad in (0, 16 - (3 * W) % 16 if (3 * W) % 16 else 32):
        got = eng.yuv420_to_bgr(torch.from_numpy(yuv).cuda(), H, W, full, pitch=3 * W + pad)
        assert np.array_equal(got.cpu().numpy()[:, :, :3 * W].reshape(n, H, W, 3), want), (H, W, full, pad)


def test_yuv_cube_on_device(torch, eng, ):
    """Every (Y, U, V) triple through the device kernel equals the oracle (both ranges)."""
    from oracle import yuv_oracle as YO
    U, V = np.meshgrid(np.arange(256, dtype=np.uint8), np.arange(256, dtype=np.uint8), indexing="ij")
    for full in (False, True):
        frames, want = [], []
        for y0 in range(64):
            Y = np.empty((512, 512), np.uint8)
            Y[0::2, 0::2], Y[0::2, 1::2], Y[1::2, 0::2], Y[1::2, 1::2] = y0, (y0 + 64) & 255, (y0 + 128) & 255, (y0 + 192) & 255
            frames.append(np.concatenate([Y.reshape(-1), U.reshape(-1), V.reshape(-1)]))
            want.append(YO.yuv420_to_bgr(Y, U, V, full))
        got = eng.yuv420_to_bgr(torch.from_numpy(np.stack(frames)).cuda(), 512, 512, full).cpu().numpy().reshape(64, 512, 512, 3)
        assert np.array_equal(got, np.stack(want)), full


@pytest.mark.parametrize("H,W,fmt", [(720, 1280, "i420"), (360, 640, "j420"), (480, 854, "i420")])
def test_host_yuv_path_equals_host_bgr_path(synth, eng, H, W, fmt):
    """avd_analyze_host_yuv420 on the decoder's pictures == avd_analyze_host on the BGR frames cv2 would have made of
    them: records byte for byte (same K1 / K2 downstream of the conversion)."""
    from oracle import yuv_oracle as YO
    clip = synth.Clip(60, H, W, seed=41, events=True, pan=2)
    bgr0 = [clip[i] for i in range(0, 60, 3)]                       # 20 samples
    Hc, Wc = H & ~1, W & ~1
    yuv = [synth.bgr_to_i420(f[:Hc, :Wc]) for f in bgr0]               # "decoder output"
    bgr = [YO.i420_to_bgr(f, Hc, Wc, fmt == "j420") for f in yuv]   # what cv2.VideoCapture.retrieve() returns for it
    a = eng.analyze_host(yuv, fmt=fmt, shape=(Hc, Wc))
    b = eng.analyze_host(bgr)
    assert a.tobytes() == b.tobytes()
    assert (a["flags"][1:] & 2).all()


def _write_clip(cv2, path, fourcc, synth, n, H, W, seed, fps=30.0):
    vw = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*fourcc), fps, (W, H))
    if not vw.isOpened():
        return False
    clip = synth.Clip(n, H, W, seed, events=True, pan=2)
    for i in range(n):
        vw.write(clip[i])
    vw.release()
    return True


@pytest.mark.parametrize("fourcc,ext", [("mp4v", "mp4"), ("MJPG", "avi")])
def test_analyze_path_yuv_route_equals_cv2_route_and_reference_body(tmp_path, synth, eng, monkeypatch, fourcc, ext):
    """video.analyze(path) on a real lossy file: the 4:2:0 route (libav pictures -> device conversion) returns exactly
    the dict of the cv2.VideoCapture route, and both agree with the reference's own cv2/NumPy body on the frames
    cv2.VideoCapture decodes.  meta fps = 2 -> step 1 -> 110 samples = three overlapped runs with halos."""
    cv2 = pytest.importorskip("cv2")
    from avd_b200 import decode, video
    from oracle import cv_body
    path = str(tmp_path / f"c.{ext}")
    if not _write_clip(cv2, path, fourcc, synth, 110, 240, 320, 51):
        pytest.skip(f"no {fourcc} encoder")
    meta = {"fps": 2.0, "width": 320, "height": 240, "duration": 55.0}
    monkeypatch.setenv("AVD_DECODE", "cv2")
    via_cv2 = video.analyze(path, dict(meta), engine=eng)
    monkeypatch.setenv("AVD_DECODE", "auto")
    src = decode.open_yuv_source(path)
    if src is None:
        pytest.skip("libav shim unavailable")
    src.release()
    via_yuv = video.analyze(path, dict(meta), engine=eng)
    assert via_yuv == via_cv2
    cap, frames = cv2.VideoCapture(path), []
    while True:
        ok, f = cap.read()
        if not ok:
            break
        frames.append(f)
    tex, fm, fv, dup, tl = cv_body.run_chunk(frames)
    assert len(via_yuv["timeline"]) == 55 and len(tl) == 110
    for a, b in zip(via_yuv["timeline"], tl[:55]):
        assert close(a, b, atol=5e-5)
    assert via_yuv["summary"]["dup_density"] == dup / 109
    assert close(via_yuv["summary"]["flow_mean"], float(np.mean(fm)), atol=summary_flow_atol(fm, "flow_mean"))
    assert close(via_yuv["summary"]["texture_var"], float(np.var(tex)), rtol=1e-9)


def test_full_1080p_clip_vs_reference_golden(synth, eng):
    """BASELINE configs[1], the WHOLE clip: 120 samples of the 1080p seed-1 clip against what the unmodified reference
    produced (tests/golden/golden_1080p_full.json): integer fields exact for every sample, flow statistics for all 119
    pairs, the returned dict, and fuse()'s verdict."""
    from avd_b200 import fusion, hints, video
    with open(os.path.join(ROOT, "tests", "golden", "golden_1080p_full.json")) as f:
        g = json.load(f)
    clip = synth.Clip(g["n_frames"], g["H"], g["W"], g["seed"], fps=g["fps"])
    frames = clip.sampled(g["step"])
    assert sha(frames[:8]) == g["input_sha_first8"]
    recs = eng.analyze_host(frames)
    assert len(recs) == 120
    assert [r["hash"].tolist() for r in recs] == g["hash_words"]
    assert [int(x) for x in recs["ham_prev"][1:]] == g["ham"]
    assert [int(x) for x in recs["lap_sum"]] == g["lap_sum"] and [int(x) for x in recs["lap_sumsq"]] == g["lap_sumsq"]
    worst = 0.0
    for r, (m, v) in zip(recs[1:], g["flow_stats"]):
        assert close(float(r["flow_mean"]), m, atol=flow_atol(m)), (float(r["flow_mean"]), m)
        assert close(float(r["flow_var"]), v, atol=flow_atol(m)), (float(r["flow_var"]), v)
        if m >= 1e-2:
            worst = max(worst, abs(float(r["flow_mean"]) - m) / m)
    assert worst <= 1e-4
    out = video.result_from_records(recs, g["H"], g["W"], g["fps"], g["meta"]["duration"], w=g["W"], h=g["H"])
    ref = g["result"]
    assert len(out["timeline"]) == len(ref["timeline"]) == 60
    for a, b in zip(out["timeline"], ref["timeline"]):
        assert close(a, b)
    means = [m for m, _ in g["flow_stats"]]
    for k, b in ref["summary"].items():
        a = out["summary"][k]
        if k in ("w", "h", "dup_density", "scene_change_rate", "fps"):
            assert a == b, k
        else:
            assert close(a, b, atol=summary_flow_atol(means, k)), (k, a, b)
    assert hints.compute_hints(dict(g["hints_meta"]), "") == g["hints"]
    fused = fusion.fuse(fusion.neutral_audio(60), copy.deepcopy(out), g["hints"], with_raw=True)
    want = g["fused_with_neutral_audio"]
    assert fused["result"]["label"] == want["result"]["label"] and fused["peaks"] == want["peaks"]
    assert abs(fused["_raw_score"] - float(np.mean(want["timeline_binned"]))) <= 0.05   # sanity: same scale
    assert abs(fused["result"]["ai_score"] - want["result"]["ai_score"]) <= 0.01 + 1e-9
    for a, b in zip(fused["timeline_binned"], want["timeline_binned"]):
        assert abs(a - b) <= 1e-3


def test_analyze_path_reuses_pinned_staging(tmp_path, synth, eng):
    """video.analyze must not allocate pinned memory per call.  Ten calls: the number of live
    cudaHostAlloc regions stops growing after the first."""
    cv2 = pytest.importorskip("cv2")
    from avd_b200 import video
    path = str(tmp_path / "c.avi")
    if not _write_clip(cv2, path, "MJPG", synth, 40, 240, 320, 61):
        pytest.skip("no MJPG encoder")
    meta = {"fps": 30.0, "width": 320, "height": 240, "duration": 40 / 30}
    first = video.analyze(path, dict(meta), engine=eng)
    n0 = eng.pinned_regions
    assert n0 >= 2
    for _ in range(9):
        assert video.analyze(path, dict(meta), engine=eng) == first
    assert eng.pinned_regions == n0
    a = eng.pinned((1024,))
    assert eng.pinned_regions == n0 + 1
    eng.free_pinned(a)
    assert eng.pinned_regions == n0


def test_cached_graph_is_recaptured_after_workspace_growth(synth, torch):
    """a cached CUDA graph holds the addresses of the library's grow-only scratch buffers; a later,
    larger call moves them.  The engine must notice (avd_ws_generation) and re-capture instead of replaying."""
    from avd_b200.engine import Engine
    e = Engine(0)
    try:
        clip = synth.Clip(64, 360, 640, seed=71, events=True)
        small = torch.from_numpy(np.stack([clip[i] for i in range(0, 16, 2)])).cuda()
        big = torch.from_numpy(np.stack([clip[i] for i in range(64)])).cuda()
        want = e.analyze_device(small).cpu().numpy().tobytes()
        assert e.analyze_device_graph(small).cpu().numpy().tobytes() == want
        gen = e.ws_generation
        e.analyze_device(big)                                   # grows the Farneback workspace, thumbnails' scratch, ...
        assert e.ws_generation > gen
        assert e.analyze_device_graph(small).cpu().numpy().tobytes() == want
        assert e.analyze_device_graph(small).cpu().numpy().tobytes() == want
        for k in range(12):                                     # the cache is bounded
            e.analyze_device_graph(small[:k % 4 + 2].clone())
        assert len(e._graphs) <= e.GRAPH_CACHE
    finally:
        e.close()


def test_engine_pool_serves_concurrent_requests(synth):
    """Requests on pool threads (api.py:133) lease separate contexts: every caller gets the dict it would get alone."""
    from avd_b200 import video
    from avd_b200.engine import Engine, engine_pool
    clips = [synth.Clip(60, 360, 640, seed=130 + i, fps=30.0) for i in range(8)]
    reqs = [(list(c.sampled(15)), {"fps": 30.0, "width": 640, "height": 360, "duration": 2.0}) for c in clips]
    solo = Engine(0)
    want = [video.analyze_frames(f, m, engine=solo) for f, m in reqs]
    solo.close()
    got, errs, gate = [None] * 8, [], threading.Barrier(8)

    def work(i):
        try:
            gate.wait()
            for _ in range(4):
                got[i] = video.analyze_frames(*reqs[i])         # engine=None: leased from the device's pool
        except Exception as e:                                  # pragma: no cover
            errs.append(e)

    ts = [threading.Thread(target=work, args=(i,)) for i in range(8)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errs and got == want
    assert 1 <= len(engine_pool(0)._all) <= engine_pool(0).size


def test_halo_frame_folded_into_the_main_launch(synth, torch, eng):
    """A chunk whose halo sample sits directly in front of it in memory goes through one K1 launch and one K2 pass
    (engine.analyze_device); the records must be those of the separate-halo path and of the unsharded run."""
    clip = synth.Clip(60, 720, 1280, seed=77, events=True, pan=2)
    allf = torch.from_numpy(np.stack([clip[i] for i in range(0, 60, 5)])).cuda()          # 12 samples
    whole = eng.analyze_device(allf).cpu().numpy().tobytes()
    l0 = eng.launches
    folded = eng.analyze_device(allf[1:], prev_frame=allf[0]).cpu().numpy().tobytes()
    n_folded = eng.launches - l0
    halo = allf[0].clone()
    l0 = eng.launches
    separate = eng.analyze_device(allf[1:], prev_frame=halo).cpu().numpy().tobytes()
    n_separate = eng.launches - l0
    assert folded == separate == whole[160:]
    assert n_folded < n_separate


def test_plain_c_client_of_the_abi(eng):
    """tests/native/abi_smoke.c on the GPU: every K1 path, the 4:2:0 route, Farneback, determinism, no pinned leak."""
    import subprocess
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "avd_b200", "csrc"), "abi_smoke"])
    out = subprocess.run([os.path.join(ROOT, "tests", "native", "abi_smoke")], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "abi_smoke ok" in out.stdout, out.stdout + out.stderr


def test_split_flow_chains_equal_single_chain_and_replay_in_a_graph(synth, torch, monkeypatch):
    """K2 runs the flow iterations of a batch of >= 32 pairs as two chains on two streams (AVD_K2_SPLIT, default 2).
    The records must be those of the single-chain run, byte for byte, also when the fork/join is captured in a CUDA graph,
    and a pair's statistics must not depend on the batch it is computed in."""
    import subprocess, sys, textwrap
    code = textwrap.dedent("""
        import sys, numpy as np, torch
        sys.path.insert(0, %r)
        import avd_b200.synth as synth
        from avd_b200.engine import Engine
        eng = Engine(0)
        clip = synth.Clip(90, 360, 640, seed=88, events=True, pan=2)
        fr = torch.from_numpy(np.stack([clip[i] for i in range(0, 90, 2)])).cuda()       # 45 samples, 44 pairs
        a = eng.analyze_device(fr).cpu().numpy().tobytes()
        g1 = eng.analyze_device_graph(fr).cpu().numpy().tobytes()
        g2 = eng.analyze_device_graph(fr).cpu().numpy().tobytes()
        small = eng.analyze_device(fr[:20]).cpu().numpy().tobytes()                       # 19 pairs: one chain
        assert a == g1 == g2 and small == a[:20 * 160]
        sys.stdout.write(__import__("hashlib").sha256(a).hexdigest())
    """ % ROOT)
    outs = []
    for split in ("1", "2"):
        env = dict(os.environ, AVD_K2_SPLIT=split)
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=600)
        assert r.returncode == 0, r.stdout + r.stderr
        outs.append(r.stdout.strip())
    assert outs[0] == outs[1] and len(outs[0]) == 64


@pytest.mark.parametrize("H,W", [(1080, 1920), (720, 1280), (480, 854), (272, 480), (33, 35), (641, 1283)])
def test_kernels_write_only_inside_caller_buffers(synth, torch, eng, H, W):
    """compute-sanitizer is closed on the GPU pool, so out-of-bounds WRITES are hunted with canaries: every output buffer of
    the device entry points (records, thumbnails, 32x32 means, BGR frames of K0) sits inside a larger allocation filled with a
    pattern; after the calls the surroundings must be untouched, and the library's own workspaces are guarded the same way
    (avd_check_guards).  Covers the streaming, banded and any-width per-frame kernels, K0 (both kernels) and K2."""
    import ctypes as C
    n, PAD = 5, 4096
    fr = torch.from_numpy(synth.noise_frames(n, H, W, 23)).cuda()

    def padded(nbytes):
        buf = torch.full((nbytes + 2 * PAD,), 0xA5, dtype=torch.uint8, device="cuda")
        return buf, buf[PAD:PAD + nbytes]

    rec_b, rec = padded(n * 160)
    th_b, th = padded(n * 320 * 320)
    ar_b, ar = padded(n * 1024)
    L, h = eng.L, eng.h
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    with eng._lock:
        eng._check(L.avd_frames_features(h, C.c_void_p(fr.data_ptr()), n, H, W, 3 * W, 3 * W * H, None, C.c_void_p(rec.data_ptr()),
                                         C.c_void_p(th.data_ptr()), C.c_void_p(ar.data_ptr()), st), "avd_frames_features")
        eng._check(L.avd_flow_stats(h, C.c_void_p(th.data_ptr()), n, None, C.c_void_p(rec.data_ptr()), st), "avd_flow_stats")
    He, We = H & ~1, W & ~1
    yuv = torch.randint(0, 256, (n, He * We * 3 // 2), dtype=torch.uint8, device="cuda")
    pitch = 3 * We + (16 - (3 * We) % 16) % 16
    bg_b, bg = padded(n * He * pitch)
    with eng._lock:
        eng._check(L.avd_yuv420_to_bgr(h, C.c_void_p(yuv.data_ptr()), yuv.stride(0), n, He, We, 0, C.c_void_p(bg.data_ptr()), pitch, He * pitch, st),
                   "avd_yuv420_to_bgr")
    torch.cuda.synchronize()
    for name, b, nbytes in (("records", rec_b, n * 160), ("thumbnails", th_b, n * 102400), ("area32", ar_b, n * 1024), ("bgr", bg_b, n * He * pitch)):
        assert bool((b[:PAD] == 0xA5).all()) and bool((b[PAD + nbytes:] == 0xA5).all()), (name, H, W)
    assert eng.check_guards() == 0


def test_config2_4k_head_vs_reference_golden(synth, eng):
    """BASELINE configs[2] geometry at full frame size: the first 12 samples of the 4K 60 fps seed-2 clip against the
    reference's own analyze() run on the whole clip (tests/golden/golden_4k_full.json): integers exact, flow in tolerance."""
    from oracle import parity
    with open(os.path.join(ROOT, "tests", "golden", "golden_4k_full.json")) as f:
        g = json.load(f)
    clip = synth.Clip(g["n_frames"], g["H"], g["W"], g["seed"], fps=g["fps"])
    frames = np.stack([clip[i * g["step"]] for i in range(12)])
    recs = eng.analyze_host(frames)
    v = parity.records_vs_golden_range(recs, g, 0, first_has_prev=False)
    assert v["ok"] and v["int_exact"], v
    # a shard that starts in the middle of the clip, seeded by its halo sample (what rank r > 0 computes)
    recs2 = eng.analyze_host(frames[4:])
    v2 = parity.records_vs_golden_range(recs2[1:], g, 5, first_has_prev=True)
    assert v2["ok"], v2


def test_config3_clip_batch_vs_reference_golden(synth, eng):
    """BASELINE configs[3]: 16 of the 256 short 720p clips, frames -> fused verdicts on the device (K1 + K2 + K3), against the
    reference's analyze() + fuse() dicts for the same clips (tests/golden/golden_clips256.json)."""
    from concurrent.futures import ThreadPoolExecutor
    from avd_b200 import fusion, hints, scoring
    from oracle import parity
    with open(os.path.join(ROOT, "tests", "golden", "golden_clips256.json")) as f:
        g = json.load(f)
    gold = {c["seed"]: c for c in g["clips"]}
    seeds = list(range(1000, 1256, 16))
    with ThreadPoolExecutor(8) as ex:
        clips = list(ex.map(lambda sd: synth.Clip(g["n_frames"], g["H"], g["W"], sd, fps=g["fps"]).sampled(g["step"]), seeds))
    meta = {"fps": 30.0, "width": 1280, "height": 720, "duration": 5.0, "bit_rate": g["hints_meta_bit_rate"]}
    hn = hints.compute_hints(dict(meta), "")
    assert hn == g["hints"]
    out = scoring.score_clip_batch(eng, clips, [dict(meta) for _ in clips], audios=[fusion.neutral_audio(5) for _ in clips],
                                   hints=[hn for _ in clips], with_raw=True)
    for sd, c, o in zip(seeds, clips, out):
        ref = gold[sd]
        assert parity.video_dict_vs_golden(o["video"], ref["video"]), (sd, o["video"], ref["video"])
        assert parity.fused_vs_golden(o["fusion"], ref["fused"]), (sd, o["fusion"], ref["fused"])


def test_yuv_entry_point_argument_errors(torch, eng):
    """The 4:2:0 entry points refuse what they cannot reproduce (odd sizes, short buffers, unknown formats) with an ordinary
    exception, and the engine stays usable."""
    from avd_b200._lib import AvdError
    from avd_b200.engine import HostBatch
    ok = np.zeros(64 * 96 * 3 // 2, np.uint8)
    assert len(eng.analyze_host([ok, ok], fmt="i420", shape=(64, 96))) == 2
    with pytest.raises(AvdError):
        eng.analyze_host([np.zeros(65 * 96 * 3 // 2, np.uint8)], fmt="i420", shape=(65, 96))      # odd height
    with pytest.raises(AvdError):
        eng.analyze_host([ok[:-1]], fmt="i420", shape=(64, 96))                                    # short frame
    with pytest.raises(AvdError):
        HostBatch([ok], fmt="nv12", shape=(64, 96))
    with pytest.raises(AvdError):
        eng.analyze_host([ok], fmt="i420")                                                         # shape missing
    with pytest.raises(AvdError):
        eng.yuv420_to_bgr(torch.zeros((1, 33 * 34 * 3 // 2), dtype=torch.uint8, device="cuda"), 33, 34)
    r = eng.analyze_host([ok, ok], fmt="j420", shape=(64, 96))
    assert int(r[1]["ham_prev"]) == 0 and eng.check_guards() == 0

"""GPU parity tests (run on the B200 box: pytest -m gpu).  Everything goes through the C ABI (libavd.so via
ctypes); the checker is the CPU oracle and the golden fixtures made from the reference's own code.

Gates (BASELINE.json north_star / SURVEY.md s8d): integer work bit-exact (gray, 32x32 area, hash bits,
Hamming, thumbnails, Laplacian sums); float features |d| <= 1e-4 |ref| + floor; fused ai_score within 1e-3.
"""
import hashlib
import os

import numpy as np
import pytest

from conftest import FLOAT_RTOL, FLOW_ATOL, close, flow_atol, summary_flow_atol

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


def _frame(synth, c):
    if c["kind"] == "noise":
        return synth.noise_frames(1, c["H"], c["W"], c["seed"])[0]
    return synth.Clip(4, c["H"], c["W"], c["seed"], events=False)[1]


def _rec_np(eng, rec):
    return eng.records_to_numpy(rec)


def test_extension_is_loaded_and_native(eng):
    from avd_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH)
    maps = open("/proc/self/maps").read()
    assert "libavd.so" in maps
    assert eng.launches == 0


@pytest.mark.parametrize("mode", [0, 1, 2, 3])
@pytest.mark.parametrize("idx", range(13))
def test_frame_features_vs_reference_golden(golden, golden_arrays, synth, torch, eng, idx, mode):
    """Every library code path (general / integer INTER_AREA, 2x2 INTER_LINEAR shortcut, upsampling, odd
    widths, 32x32 minimum, 4K) against what cv2 produced for the reference; all three K1 implementations
    (0 = streaming fused kernel where the shape allows, 1 = generic, 2 = banded fused kernel, 3 = any-width fused kernel)."""
    c = golden["frames"][idx]
    f = _frame(synth, c)
    eng.set_k1_mode(mode)
    try:
        d = torch.from_numpy(f).cuda()
        assert sha(eng.gray(d).cpu().numpy()) == c["gray_sha"]
        rec, thumbs, area = eng.frames_features(d[None], want_area32=True)
        r = _rec_np(eng, rec)[0]
        assert np.array_equal(area[0].cpu().numpy(), golden_arrays[c["name"] + ".area32"])
        assert r["hash"].tolist() == c["hash_words"]
        assert sha(thumbs[0].cpu().numpy()) == c["thumb_sha"]
        assert (int(r["lap_sum"]), int(r["lap_sumsq"])) == (c["lap_sum"], c["lap_sumsq"])
        assert int(r["ham_prev"]) == -1 and int(r["flags"]) == 0
    finally:
        eng.set_k1_mode(0)


def test_fused_equals_generic_on_batches(synth, torch, eng):
    """The three K1 implementations must agree byte for byte on a batch (incl. Hamming chain)."""
    # streaming kernel: pair-sum thumbnail (1280/1920/2560/3840 wide; static lanes at 1280 and 2560) and general two-tap
    # thumbnail (1344, 2048, 4096 wide), integer and fractional INTER_AREA rows; banded kernel: 128..704 wide; generic: rest
    for (H, W) in [(720, 1280), (1080, 1920), (96, 128), (270, 480), (360, 640), (960, 1280), (1440, 2560), (2160, 3840),
                   (648, 704), (1000, 4096), (648, 1344), (800, 2048), (4096, 1280), (641, 3200)]:
        clip = synth.Clip(40, H, W, seed=5, events=True)
        fr = torch.from_numpy(np.stack([clip[i] for i in (0, 3, 12, 13, 20, 25, 31, 39)])).cuda()
        outs = []
        for mode in (0, 1, 2, 3):
            eng.set_k1_mode(mode)
            rec, th, ar = eng.frames_features(fr, want_area32=True)
            outs.append((rec.cpu().numpy().tobytes(), th.cpu().numpy().tobytes(), ar.cpu().numpy().tobytes()))
        eng.set_k1_mode(0)
        assert outs[0] == outs[1], (H, W, "streaming vs generic")
        assert outs[2] == outs[1], (H, W, "banded vs generic")
        assert outs[3] == outs[1], (H, W, "any-width fused vs generic")


def test_any_width_kernel_equals_generic_on_ragged_shapes(synth, torch, eng, oracle):
    """k1_rows.cu: widths that are not a multiple of 16 (854x480 of BASELINE configs[4], 1366x768, odd crops), the 2x2
    thumbnail shortcut, thumbnail upsampling, the 32x32 minimum, unaligned views -- byte for byte the generic path, and the
    oracle on one frame per shape."""
    for (H, W) in [(480, 854), (768, 1366), (333, 777), (144, 256), (640, 640), (32, 32), (33, 35), (641, 1283), (1080, 1918), (100, 150)]:
        clip = synth.Clip(12, H, W, seed=15, events=False, pan=3)
        host = np.stack([clip[i] for i in (0, 1, 5, 5, 11)])
        fr = torch.from_numpy(host).cuda()
        outs = {}
        for mode in (1, 3):
            eng.set_k1_mode(mode)
            rec, th, ar = eng.frames_features(fr, want_area32=True)
            outs[mode] = (rec.cpu().numpy().tobytes(), th.cpu().numpy().tobytes(), ar.cpu().numpy().tobytes())
        eng.set_k1_mode(0)
        assert outs[3] == outs[1], (H, W)
        rec, th = eng.frames_features(fr)                               # auto: what a caller gets
        assert (rec.cpu().numpy().tobytes(), th.cpu().numpy().tobytes()) == outs[1][:2], (H, W)
        g = oracle.gray(host[1])
        r = _rec_np(eng, rec)[1]
        assert (int(r["lap_sum"]), int(r["lap_sumsq"])) == oracle.lap_sums(g)
        assert np.array_equal(r["hash"], oracle.hash_words(oracle.area32(g))) and np.array_equal(th[1].cpu().numpy(), oracle.thumb320(g))
    # a view that is not even 4-byte aligned (byte path of the gray stage)
    H, W = 120, 203
    f = synth.noise_frames(2, H, W, 19)
    big = torch.zeros((2, H, W * 3 + 7), dtype=torch.uint8, device="cuda")
    big[:, :, 3:3 + W * 3] = torch.from_numpy(f.reshape(2, H, W * 3)).cuda()
    view = big[:, :, 3:3 + W * 3].unflatten(2, (W, 3))
    outs = {}
    for mode in (1, 3):
        eng.set_k1_mode(mode)
        rec, th, ar = eng.frames_features(view, want_area32=True)
        outs[mode] = (rec.cpu().numpy().tobytes(), th.cpu().numpy().tobytes(), ar.cpu().numpy().tobytes())
    eng.set_k1_mode(0)
    assert outs[3] == outs[1]


def test_area_rounding_boundaries(torch, eng, oracle):
    """INTER_AREA cells whose exact mean sits on or next to a rounding boundary (k + 0.5): the streaming kernel
    must fall back to the library's float32 summation order for them and still be bit-exact with the oracle."""
    rng = np.random.default_rng(11)
    H, W = 1080, 1920
    frames = []
    g = np.full((H, W), 100, np.uint8); g[:, 1::2] = 101; frames.append(g)               # every cell mean = 100.5
    g = np.full((H, W), 7, np.uint8); g[::2, :] = 8; frames.append(g)                     # row stripes, fractional rows
    g = (100 + (rng.random((H, W)) < 0.5)).astype(np.uint8); frames.append(g)             # means scattered around 100.5
    g = rng.integers(0, 256, (H, W), dtype=np.uint8); g[:540] = (g[:540] & 1) + 200; frames.append(g)
    g = np.full((H, W), 255, np.uint8); g[:, ::2] = 254; frames.append(g)                 # 254.5: saturation side
    bgr = np.stack([np.repeat(x[:, :, None], 3, axis=2) for x in frames])                 # B=G=R=v -> gray = v
    d = torch.from_numpy(bgr).cuda()
    outs = {}
    for mode in (0, 1):
        eng.set_k1_mode(mode)
        rec, th, ar = eng.frames_features(d, want_area32=True)
        outs[mode] = (rec.cpu().numpy().tobytes(), ar.cpu().numpy())
    eng.set_k1_mode(0)
    for i, x in enumerate(frames):
        assert np.array_equal(eng.gray(d[i]).cpu().numpy(), x)
        assert np.array_equal(outs[0][1][i], oracle.area32(x)), i
    assert outs[0][0] == outs[1][0] and np.array_equal(outs[0][1], outs[1][1])


def test_streaming_kernel_two_frames_per_cta(synth, torch, eng):
    """W = 1280: a CTA of the streaming kernel takes band b of two consecutive frames side by side (80 strips per frame fill
    2.5 warps).  Odd and even frame counts (the last CTA row holds one frame), integer and fractional INTER_AREA rows."""
    for H in (720, 641, 960):
        clip = synth.Clip(40, H, 1280, seed=7, events=True)
        for n in (1, 3, 8):
            fr = torch.from_numpy(np.stack([clip[i] for i in (0, 3, 12, 13, 20, 25, 31, 39)[:n]])).cuda()
            outs = []
            for mode in (0, 1):
                eng.set_k1_mode(mode)
                rec, th, ar = eng.frames_features(fr, want_area32=True)
                outs.append((rec.cpu().numpy().tobytes(), th.cpu().numpy().tobytes(), ar.cpu().numpy().tobytes()))
            eng.set_k1_mode(0)
            assert outs[0] == outs[1], (H, n)


def test_any_width_kernel_extreme_aspect_ratios(torch, eng):
    """k1_rows.cu at the corners of its shape range: 128-row bands of two 16-pixel groups, one-row bands 4096 pixels wide,
    widths one pixel off the streaming kernel's grid -- byte for byte the generic path."""
    rng = np.random.default_rng(3)
    for (H, W) in [(4096, 33), (35, 4096), (2048, 100), (720, 1281), (719, 1279), (32, 4096), (4096, 32)]:
        fr = torch.from_numpy(rng.integers(0, 256, (3, H, W, 3), dtype=np.uint8)).cuda()
        outs = []
        for mode in (3, 1):
            eng.set_k1_mode(mode)
            rec, th, ar = eng.frames_features(fr, want_area32=True)
            outs.append((rec.cpu().numpy().tobytes(), th.cpu().numpy().tobytes(), ar.cpu().numpy().tobytes()))
        eng.set_k1_mode(0)
        assert outs[0] == outs[1], (H, W)


def test_any_width_kernel_area_rounding_boundaries(torch, eng, oracle):
    """k1_rows.cu computes INTER_AREA cells as exact integers and lists the cells whose mean is within the float32 summation
    error of a rounding boundary for area_fix_kernel (the library's order): frames built so that many / all cells sit on or
    next to k + 0.5, on ragged, fractional-scale and integer-scale shapes, against the generic path and the oracle."""
    rng = np.random.default_rng(12)
    for (H, W) in [(480, 854), (768, 1366), (480, 640), (333, 777), (600, 1920)]:
        frames = []
        g = np.full((H, W), 100, np.uint8); g[:, 1::2] = 101; frames.append(g)               # column stripes: means near 100.5
        g = np.full((H, W), 7, np.uint8); g[::2, :] = 8; frames.append(g)                     # row stripes, fractional rows
        g = (100 + (rng.random((H, W)) < 0.5)).astype(np.uint8); frames.append(g)             # means scattered around 100.5
        g = np.full((H, W), 255, np.uint8); g[:, ::2] = 254; frames.append(g)                 # 254.5: saturation side
        g = rng.integers(0, 256, (H, W), dtype=np.uint8); frames.append(g)
        bgr = np.stack([np.repeat(x[:, :, None], 3, axis=2) for x in frames])                 # B=G=R=v -> gray = v
        d = torch.from_numpy(bgr).cuda()
        outs = {}
        for mode in (3, 1):
            eng.set_k1_mode(mode)
            rec, th, ar = eng.frames_features(d, want_area32=True)
            outs[mode] = (rec.cpu().numpy().tobytes(), ar.cpu().numpy())
        eng.set_k1_mode(0)
        for i, x in enumerate(frames):
            assert np.array_equal(outs[3][1][i], oracle.area32(x)), (H, W, i)
        assert outs[3][0] == outs[1][0] and np.array_equal(outs[3][1], outs[1][1]), (H, W)


def test_streaming_kernel_padded_pitch(synth, torch, eng):
    """Row pitch > 3W (16-byte aligned) and a frame stride > H*pitch through the streaming kernel."""
    H, W = 720, 1280
    f = synth.noise_frames(3, H, W, 13)
    big = torch.zeros((3, H + 2, W * 3 + 64), dtype=torch.uint8, device="cuda")
    big[:, :H, :W * 3] = torch.from_numpy(f.reshape(3, H, W * 3)).cuda()
    view = big[:, :H, :W * 3].unflatten(2, (W, 3))
    outs = []
    for mode in (0, 1):
        eng.set_k1_mode(mode)
        rec, th, ar = eng.frames_features(view, want_area32=True)
        outs.append((rec.cpu().numpy().tobytes(), th.cpu().numpy().tobytes(), ar.cpu().numpy().tobytes()))
    eng.set_k1_mode(0)
    assert outs[0] == outs[1]


def test_padded_pitch_and_unaligned_views(synth, torch, eng, oracle):
    """Row pitch > 3W and a base pointer that is not 16-byte aligned (ragged inputs)."""
    H, W = 100, 150
    f = synth.noise_frames(2, H, W, 9)
    big = torch.zeros((2, H, W * 3 + 37), dtype=torch.uint8, device="cuda")
    big[:, :, 5:5 + W * 3] = torch.from_numpy(f.reshape(2, H, W * 3)).cuda()
    view = big[:, :, 5:5 + W * 3].unflatten(2, (W, 3))
    rec, th = eng.frames_features(view)
    r = _rec_np(eng, rec)
    for i in range(2):
        g = oracle.gray(f[i])
        assert (int(r[i]["lap_sum"]), int(r[i]["lap_sumsq"])) == oracle.lap_sums(g)
        assert np.array_equal(r[i]["hash"], oracle.hash_words(oracle.area32(g)))
        assert np.array_equal(th[i].cpu().numpy(), oracle.thumb320(g))


@pytest.mark.parametrize("name", ["shift4", "shift23", "static", "cut", "noise", "static_noise", "flat"])
def test_farneback_vs_reference_golden(golden, golden_arrays, torch, eng, oracle, name):
    c = next(p for p in golden["flow_pairs"] if p["name"] == name)
    p, n = golden_arrays[f"pair_{name}.prev"], golden_arrays[f"pair_{name}.next"]
    fl = eng.flow_fields(torch.from_numpy(p)[None].cuda(), torch.from_numpy(n)[None].cuda())[0].cpu().numpy()
    m, v = oracle.mag_stats(fl)                      # statistics of OUR field with NumPy's float32 pairwise sums
    assert close(m, c["mean"], atol=flow_atol(c["mean"])), (m, c["mean"])
    assert close(v, c["var"], atol=flow_atol(c["mean"])), (v, c["var"])
    if f"pair_{name}.flow" in golden_arrays:         # dense field against cv2's
        ref = golden_arrays[f"pair_{name}.flow"]
        assert np.abs(fl - ref).max() <= 1e-3 * max(1.0, c["flow_abs_max"])
    o = oracle.farneback(p, n)                       # and against the C restatement
    if name not in ("static",):                      # static pairs are chaotic at the last row/column
        assert np.abs(fl - o).max() <= 2e-3 * max(1.0, c["flow_abs_max"])


def test_flow_stats_kernel_matches_field_statistics(golden_arrays, torch, eng, oracle):
    """K2's own mean/var (float64 accumulation) vs NumPy float32 pairwise on the same field; chain API."""
    names = ["shift4", "cut", "noise", "static_noise"]
    thumbs = torch.from_numpy(np.stack([golden_arrays["pair_shift4.prev"]] + [golden_arrays[f"pair_{n}.next"] for n in names])).cuda()
    rec = torch.zeros((len(names) + 1, 160), dtype=torch.uint8, device="cuda")
    eng.flow_stats(thumbs, rec)
    r = _rec_np(eng, rec)
    fields = eng.flow_fields(thumbs[:-1].contiguous(), thumbs[1:].contiguous()).cpu().numpy()
    assert int(r[0]["flags"]) & 2 == 0
    for i in range(len(names)):
        m, v = oracle.mag_stats(fields[i])
        assert int(r[i + 1]["flags"]) & 2
        assert close(float(r[i + 1]["flow_mean"]), m, rtol=1e-6, atol=1e-7)
        assert close(float(r[i + 1]["flow_var"]), v, rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("name", ["clip_small", "clip_odd", "clip_nometa", "clip_720p", "clip_1080p_head"])
def test_clip_vs_reference_analyze(golden, synth, eng, name):
    """Whole path (host entry point -> records -> tail) against the unmodified reference's analyze()."""
    from avd_b200 import video
    c = golden["clips"][name]
    total = c.get("of_clip_frames", c["n_frames"])
    clip = synth.Clip(total, c["H"], c["W"], c["seed"], fps=c["fps"])
    step = video.sample_step(c["meta"].get("fps") or c["fps"])
    frames = np.stack([clip[i] for i in range(0, c["n_frames"], step)])
    assert sha(frames) == c["input_sha"]
    recs = eng.analyze_host(frames)
    assert [r["hash"].tolist() for r in recs] == c["hash_words"]
    if "ham" in c:
        assert [int(x) for x in recs["ham_prev"][1:]] == c["ham"]
    assert [int(x) for x in recs["lap_sum"]] == c["lap_sum"] and [int(x) for x in recs["lap_sumsq"]] == c["lap_sumsq"]
    for r, (m, v) in zip(recs[1:], c["flow_stats"]):
        assert close(float(r["flow_mean"]), m, atol=flow_atol(m)) and close(float(r["flow_var"]), v, atol=flow_atol(m)), (float(r["flow_mean"]), m, float(r["flow_var"]), v)
    meta = dict(c["meta"]) if c["meta"] else {"fps": c["fps"], "duration": c["n_frames"] / c["fps"]}
    out = video.analyze_frames(frames, meta, engine=eng)
    ref = c["result"]
    assert out["timeline"] is out["timeline_ai"] and len(out["timeline"]) == len(ref["timeline"])
    for a, b in zip(out["timeline"], ref["timeline"]):
        assert close(a, b)
    for k, b in ref["summary"].items():
        a = out["summary"][k]
        if k in ("w", "h", "dup_density", "scene_change_rate", "fps"):
            assert a == b, k
        else:
            assert close(a, b, atol=summary_flow_atol([m for m, _ in c["flow_stats"]], k)), (k, a, b)


def test_fused_score_and_label_vs_reference(golden, synth, eng):
    """ai_score / label after fusion: the reference's fuse() output for its own video result vs ours."""
    from avd_b200 import fusion, hints, video
    c = golden["clips"]["clip_720p"]
    clip = synth.Clip(c["n_frames"], c["H"], c["W"], c["seed"], fps=c["fps"])
    frames = clip.sampled(c["step"])
    ours = video.analyze_frames(frames, dict(c["meta"]), engine=eng)
    meta = dict(c["meta"], bit_rate=2_500_000)
    h = hints.compute_hints(meta, "")
    import copy
    a = fusion.fuse(fusion.neutral_audio(10), copy.deepcopy(ours), h, with_raw=True)
    b = fusion.fuse(fusion.neutral_audio(10), copy.deepcopy(c["result"]), h, with_raw=True)
    assert abs(a["_raw_score"] - b["_raw_score"]) <= 1e-3
    assert a["result"]["label"] == b["result"]["label"] and a["peaks"] == b["peaks"]
    assert len(a["timeline_binned"]) == len(b["timeline_binned"])


def test_decode_seam_ffv1_file(golden, synth, eng, tmp_path):
    """analyze(path, meta): host decode with cv2.VideoCapture like the reference, lossless FFV1 clip."""
    cv2 = pytest.importorskip("cv2")
    from avd_b200 import video
    c = golden["clips"]["clip_720p"]
    clip = synth.Clip(c["n_frames"], c["H"], c["W"], c["seed"], fps=c["fps"])
    path = str(tmp_path / "c.avi")
    vw = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"FFV1"), c["fps"], (c["W"], c["H"]))
    if not vw.isOpened():
        pytest.skip("no FFV1 encoder in this cv2 build")
    for i in range(c["n_frames"]):
        vw.write(clip[i])
    vw.release()
    for meta, key in ((dict(c["meta"]), "result_via_ffv1_file"), ({}, "result_via_ffv1_file_nometa")):
        out, ref = video.analyze(path, meta, engine=eng), c[key]
        assert len(out["timeline"]) == len(ref["timeline"])
        assert all(close(a, b) for a, b in zip(out["timeline"], ref["timeline"]))
        for k, b in ref["summary"].items():
            assert close(float(out["summary"][k]), float(b), atol=summary_flow_atol([m for m, _ in c["flow_stats"]], k)), k
    assert video.analyze(str(tmp_path / "missing.avi"), {}, engine=eng) == {"timeline": [], "summary": {}, "timeline_ai": []}


def test_chunk_halo_equals_single_run(synth, eng):
    """Frame-chunk sharding changes no arithmetic: records of [halo + chunk] runs == one run, byte for byte."""
    from avd_b200 import dist
    clip = synth.Clip(200, 180, 320, seed=8)
    frames = clip.sampled(10)                       # 20 samples
    whole = eng.analyze_host(frames)
    for world in (2, 3, 8):
        parts = []
        for r in range(world):
            lo, hi, halo = dist.local_range_with_halo(len(frames), r, world)
            batch = ([frames[halo]] if halo is not None else []) + [frames[i] for i in range(lo, hi)]
            rec = eng.analyze_host(batch) if batch else whole[:0]
            parts.append(rec[1:] if halo is not None else rec)
        assert np.concatenate(parts).tobytes() == whole.tobytes(), world


def test_device_path_equals_host_path(synth, torch, eng):
    clip = synth.Clip(120, 360, 640, seed=6)
    frames = clip.sampled(10)
    host = eng.analyze_host(frames)
    dev = eng.records_to_numpy(eng.analyze_device(torch.from_numpy(frames).cuda()))
    assert host.tobytes() == dev.tobytes()
    # with a halo frame on the device path
    dev2 = eng.records_to_numpy(eng.analyze_device(torch.from_numpy(frames[5:]).cuda(), prev_frame=torch.from_numpy(frames[4]).cuda()))
    assert dev2.tobytes() == host[5:].tobytes()


def test_full_size_1080p_against_oracle(synth, torch, eng, oracle):
    """BASELINE configs[1] geometry (1080p, texture clip with all events), a 24-sample slice at full size:
    integer features against the oracle for every frame, flow statistics for every pair."""
    clip = synth.Clip(1800, 1080, 1920, seed=1, fps=30.0)
    idx = list(range(480, 840, 15))                 # crosses the frozen segment and the scene cut
    frames = np.stack([clip[i] for i in idx])
    recs = eng.analyze_host(frames)
    ref = oracle.frames_features(frames)
    for r, o in zip(recs, ref):
        assert (int(r["lap_sum"]), int(r["lap_sumsq"])) == (o["lap_sum"], o["lap_sumsq"])
        assert np.array_equal(r["hash"], o["hash"]) and int(r["ham_prev"]) == o["ham"]
        if "flow_mean" in o:
            assert close(float(r["flow_mean"]), o["flow_mean"], atol=flow_atol(o["flow_mean"])), (float(r["flow_mean"]), o["flow_mean"])
            assert close(float(r["flow_var"]), o["flow_var"], atol=flow_atol(o["flow_mean"]))
    assert (recs["ham_prev"] == 0).sum() >= 2        # the frozen run is seen as duplicates


def test_size_independent_properties_at_full_size(synth, torch, eng):
    """Properties that need no oracle: duplicates -> Hamming 0 and identical records; flat frames -> zero
    Laplacian; sum of the Laplacian depends only on border rows/columns (REFLECT_101 telescoping)."""
    H, W = 2160, 3840
    f = synth.noise_frames(2, H, W, 17)
    batch = torch.from_numpy(np.stack([f[0], f[0], f[1], np.full((H, W, 3), 77, np.uint8)])).cuda()
    rec, th = eng.frames_features(batch)
    r = eng.records_to_numpy(rec)
    assert int(r[1]["ham_prev"]) == 0 and r[0]["hash"].tolist() == r[1]["hash"].tolist()
    assert (int(r[0]["lap_sum"]), int(r[0]["lap_sumsq"])) == (int(r[1]["lap_sum"]), int(r[1]["lap_sumsq"]))
    assert torch.equal(th[0], th[1])
    assert int(r[3]["lap_sum"]) == 0 and int(r[3]["lap_sumsq"]) == 0 and bool((th[3] == 77).all())
    g = eng.gray(batch[2]).cpu().numpy().astype(np.int64)
    want = (g[1].sum() + g[-2].sum() - g[0].sum() - g[-1].sum()) + (g[:, 1].sum() + g[:, -2].sum() - g[:, 0].sum() - g[:, -1].sum())
    assert int(r[2]["lap_sum"]) == int(want)


def test_edge_cases_and_errors(torch, eng):
    from avd_b200._lib import AvdError
    assert len(eng.analyze_host([])) == 0
    one = eng.analyze_host([np.zeros((32, 32, 3), np.uint8)])         # minimum size, single frame
    assert int(one[0]["ham_prev"]) == -1 and int(one[0]["flags"]) == 0 and float(one[0]["flow_mean"]) == 0.0
    with pytest.raises(AvdError):
        eng.analyze_host([np.zeros((16, 64, 3), np.uint8)])            # H < 32: INTER_AREA upsampling is off-path
    with pytest.raises(AvdError):
        eng.analyze_host([np.zeros((64, 64, 3), np.uint8), np.zeros((64, 65, 3), np.uint8)])
    assert len(eng.analyze_host([np.zeros((40, 50, 3), np.uint8)])) == 1   # engine still usable after errors


@pytest.mark.parametrize("H,W", [(720, 1280), (480, 854)])
def test_device_path_graph_replay_equals_plain_launches(synth, torch, eng, H, W):
    """Engine.analyze_device_graph (CUDA graph of the step's launches) must return the same records as the plain
    launches, on first use (capture) and on replays with new frame contents at the same addresses, with and
    without a chunk halo; through the streaming kernel (two frames per CTA at this width) and the any-width kernel
    (both append to the fix list inside the captured graph)."""
    clip = synth.Clip(40, H, W, seed=21, events=True)
    frames = torch.from_numpy(np.stack([clip[i] for i in range(0, 40, 4)])).cuda()
    halo = torch.from_numpy(clip[39]).cuda()
    for prev in (None, halo):
        for rnd in range(3):
            if rnd:
                frames.copy_(torch.from_numpy(np.stack([clip[(i + 7 * rnd) % 40] for i in range(0, 40, 4)])))
            want = eng.analyze_device(frames, prev_frame=prev).cpu().numpy().tobytes()
            got = eng.analyze_device_graph(frames, prev_frame=prev).cpu().numpy().tobytes()
            assert got == want, (prev is not None, rnd)


def test_concurrent_calls_from_pool_threads(synth, eng):
    """api.py:133 runs the analyzer on pool threads, so several requests can be inside the same process-wide engine at
    once: calls are serialised by the engine's lock and every thread gets the records of its own frames."""
    import threading
    clips = [synth.Clip(24, 360, 640, seed=90 + i, events=False) for i in range(4)]
    batches = [[c[j] for j in (0, 8, 16, 23)] for c in clips]
    want = [eng.analyze_host(b).tobytes() for b in batches]
    got, errs = [None] * 4, []

    def work(i):
        try:
            for _ in range(5):
                got[i] = eng.analyze_host(batches[i]).tobytes()
        except Exception as e:                      # pragma: no cover
            errs.append(e)

    ts = [threading.Thread(target=work, args=(i,)) for i in range(4)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errs and got == want


def test_request_batcher_on_device(synth, eng):
    """service.RequestBatcher with the real engine: eight requests released together, every caller gets the dict it
    would have got alone, in fewer engine runs."""
    import threading
    from avd_b200 import service, video
    clips = [synth.Clip(60, 360, 640, seed=120 + i, fps=30.0) for i in range(8)]
    reqs = [(list(c.sampled(15)), {"fps": 30.0, "width": 640, "height": 360, "duration": 2.0}) for c in clips]
    want = [video.analyze_frames(f, m, engine=eng) for f, m in reqs]
    b = service.RequestBatcher(engine=eng, max_wait_ms=200.0)
    got, gate = [None] * 8, threading.Barrier(8)

    def work(i):
        gate.wait()
        got[i] = b.analyze_frames(*reqs[i])

    ts = [threading.Thread(target=work, args=(i,)) for i in range(8)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert got == want
    assert sum(b.batches) == 8 and len(b.batches) < 8

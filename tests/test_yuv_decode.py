"""CPU tests of the decode -> device seam (first step): the oracle of the decoder-side colour
conversion is pinned to libswscale's own output (tests/golden/golden_yuv.json, made by calling sws_scale of the
library cv2 ships) and to cv2.VideoCapture on real files; the libav shim delivers the same pictures cv2 decodes."""
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import ROOT


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def golden_yuv():
    with open(os.path.join(ROOT, "tests", "golden", "golden_yuv.json")) as f:
        return json.load(f)


def _planes(seed, H, W):
    rng = np.random.default_rng(seed)
    return (rng.integers(0, 256, (H, W), dtype=np.uint8), rng.integers(0, 256, (H // 2, W // 2), dtype=np.uint8),
            rng.integers(0, 256, (H // 2, W // 2), dtype=np.uint8))


def _cube(y0):
    U, V = np.meshgrid(np.arange(256, dtype=np.uint8), np.arange(256, dtype=np.uint8), indexing="ij")
    Y = np.empty((512, 512), np.uint8)
    Y[0::2, 0::2], Y[0::2, 1::2], Y[1::2, 0::2], Y[1::2, 1::2] = y0, (y0 + 64) & 255, (y0 + 128) & 255, (y0 + 192) & 255
    return Y, np.ascontiguousarray(U), np.ascontiguousarray(V)


def test_oracle_conversion_equals_libswscale_on_random_planes(golden_yuv):
    from oracle import yuv_oracle as YO
    cases = [c for c in golden_yuv["cases"] if c["kind"] == "random"]
    assert len(cases) == 14
    for c in cases:
        Y, U, V = _planes(c["seed"], c["H"], c["W"])
        bgr = YO.yuv420_to_bgr(Y, U, V, c["full_range"])
        assert bgr[:4, :4].tolist() == c["corner"], c
        assert sha(bgr) == c["sha256"], (c["H"], c["W"], c["full_range"])


def test_oracle_conversion_equals_libswscale_on_the_whole_yuv_cube(golden_yuv):
    """Every (Y, U, V) triple, both ranges: 256 luma values x 65536 chroma pairs."""
    from oracle import yuv_oracle as YO
    cases = [c for c in golden_yuv["cases"] if c["kind"] == "cube"]
    assert len(cases) == 128
    for c in cases:
        assert sha(YO.yuv420_to_bgr(*_cube(c["y0"]), c["full_range"])) == c["sha256"], c


def test_conversion_coefficients():
    from oracle import yuv_oracle as YO
    assert YO.coefficients(False) == {"y": 9539, "vr": 13075, "ub": 16525, "vg": -6660, "ug": -3209, "yo": 128}
    assert YO.coefficients(True) == {"y": 8192, "vr": 11485, "ub": 14516, "vg": -5850, "ug": -2819, "yo": 0}


def _write_clip(cv2, path, fourcc, synth, n=24, H=240, W=320, seed=31):
    vw = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*fourcc), 30.0, (W, H))
    if not vw.isOpened():
        return False
    clip = synth.Clip(n, H, W, seed, events=False, pan=3)
    for i in range(n):
        vw.write(clip[i])
    vw.release()
    return True


@pytest.mark.parametrize("fourcc,ext,fmt", [("mp4v", "mp4", "i420"), ("MJPG", "avi", "j420")])
def test_libav_shim_plus_oracle_conversion_equals_cv2_videocapture(tmp_path, synth, fourcc, ext, fmt):
    """The decoder's 4:2:0 pictures (read through libav before any conversion) converted by the oracle are, byte for
    byte, the BGR frames cv2.VideoCapture returns for the same file -- MPEG-4 (limited range) and MJPEG (full range)."""
    cv2 = pytest.importorskip("cv2")
    from avd_b200 import decode
    from oracle import yuv_oracle as YO
    path = str(tmp_path / f"c.{ext}")
    if not _write_clip(cv2, path, fourcc, synth):
        pytest.skip(f"no {fourcc} encoder in this cv2 build")
    try:
        decode.load_libav()
    except decode.DecodeUnavailable as e:
        pytest.skip(str(e))
    src = decode.open_yuv_source(path)
    assert src is not None
    cap = cv2.VideoCapture(path)
    n = 0
    while True:
        a, b = src.grab(), cap.grab()
        assert a == b
        if not a:
            break
        ok, bgr = cap.retrieve()
        assert ok and src.fmt == fmt and (src.height, src.width) == bgr.shape[:2]
        buf = np.empty(src.frame_nbytes(), np.uint8)
        assert src.retrieve_into(buf)
        assert np.array_equal(YO.i420_to_bgr(buf, src.height, src.width, fmt == "j420"), bgr), n
        n += 1
    assert n == 24
    src.release()
    cap.release()


def test_yuv_route_refuses_what_it_cannot_reproduce(tmp_path, synth):
    """Streams that are not 8-bit 4:2:0 (here: FFV1, which OpenCV writes as bgra/bgr0) stay on cv2.VideoCapture."""
    cv2 = pytest.importorskip("cv2")
    from avd_b200 import decode
    path = str(tmp_path / "c.avi")
    if not _write_clip(cv2, path, "FFV1", synth, n=3):
        pytest.skip("no FFV1 encoder")
    src = decode.open_yuv_source(path)
    if src is not None:
        with pytest.raises(decode.DecodeUnavailable):
            src.grab()
        src.release()
    assert decode.open_yuv_source(str(tmp_path / "missing.mp4")) is None


def test_oracle_reproduces_full_1080p_reference_golden_slice(oracle, synth):
    """BASELINE configs[1] at full size: 16 consecutive samples of the 120 (crossing the frozen run) against what the
    unmodified reference produced for the whole clip (tests/golden/golden_1080p_full.json)."""
    from conftest import close, flow_atol
    with open(os.path.join(ROOT, "tests", "golden", "golden_1080p_full.json")) as f:
        g = json.load(f)
    clip = synth.Clip(g["n_frames"], g["H"], g["W"], g["seed"], fps=g["fps"])
    lo, hi = 30, 46
    frames = [clip[i * g["step"]] for i in range(lo, hi)]
    recs = oracle.frames_features(frames)
    for k, r in enumerate(recs):
        i = lo + k
        assert r["hash"].tolist() == g["hash_words"][i]
        assert (r["lap_sum"], r["lap_sumsq"]) == (g["lap_sum"][i], g["lap_sumsq"][i])
        assert sha(r["thumb"]) == g["thumb_sha"][i]
        if k:
            m, v = g["flow_stats"][i - 1]
            assert r["ham"] == g["ham"][i - 1]
            assert close(r["flow_mean"], m, atol=flow_atol(m)) and close(r["flow_var"], v, atol=flow_atol(m)), (i, r["flow_mean"], m)


def test_oracle_reproduces_configs3_clip_golden(oracle, synth):
    """One of the 256 configs[3] clips: the oracle's records -> tail -> fuse against the reference's dicts
    (tests/golden/golden_clips256.json)."""
    from oracle import parity
    with open(os.path.join(ROOT, "tests", "golden", "golden_clips256.json")) as f:
        g = json.load(f)
    ref = next(c for c in g["clips"] if c["seed"] == 1003)
    frames = synth.Clip(g["n_frames"], g["H"], g["W"], 1003, fps=g["fps"]).sampled(g["step"])
    recs = oracle.frames_features(frames)
    assert recs[0]["lap_sum"] == ref["lap_sum0"] and recs[0]["hash"].tolist() == ref["hash0"]
    v = oracle.video_result(recs, g["H"], g["W"], g["fps"], 5.0)
    assert parity.video_dict_vs_golden(v, ref["video"])
    f = oracle.fuse({"scores": {}, "flags_audio": {}, "timeline": [0.5] * 5}, v, g["hints"])
    assert parity.fused_vs_golden(f, ref["fused"])

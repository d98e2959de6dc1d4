"""CPU: pin the oracle (oracle/) against the goldens produced by the reference's own code
(tests/golden/make_golden.py).  Integer stages bit-exact; Farneback statistics within tolerance."""
import hashlib

import numpy as np
import pytest

from conftest import FLOW_ATOL, close, flow_atol, summary_flow_atol


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _frame(synth, c):
    if c["kind"] == "noise":
        return synth.noise_frames(1, c["H"], c["W"], c["seed"])[0]
    return synth.Clip(4, c["H"], c["W"], c["seed"], events=False)[1]


def test_generator_is_stable(golden, synth):
    for c in golden["frames"]:
        assert sha(_frame(synth, c)) == c["input_sha"], c["name"]


@pytest.mark.parametrize("idx", range(13))
def test_frame_primitives(golden, golden_arrays, oracle, synth, idx):
    c = golden["frames"][idx]
    f = _frame(synth, c)
    g = oracle.gray(f)
    assert sha(g) == c["gray_sha"]
    a = oracle.area32(g)
    assert np.array_equal(a, golden_arrays[c["name"] + ".area32"])
    assert oracle.hash_words(a).tolist() == c["hash_words"]
    t = oracle.thumb320(g)
    assert sha(t) == c["thumb_sha"]
    if c["name"] + ".thumb" in golden_arrays:
        assert np.array_equal(t, golden_arrays[c["name"] + ".thumb"])
    s, ss = oracle.lap_sums(g)
    assert (s, ss) == (c["lap_sum"], c["lap_sumsq"])
    assert abs(oracle.texture_from_sums(s, ss, c["H"] * c["W"]) - c["lap_var"]) <= 1e-12 * abs(c["lap_var"])


def test_gray_exhaustive_formula(oracle):
    # all 2^24 colours against the closed form of A.1 (vectorised); cv2 itself was probed at survey time
    b, g, r = np.meshgrid(np.arange(256, dtype=np.uint32), np.arange(256, dtype=np.uint32), np.arange(0, 256, 5, dtype=np.uint32), indexing="ij")
    img = np.stack([b, g, r], -1).astype(np.uint8).reshape(256, -1, 3)
    want = ((3735 * b + 19235 * g + 9798 * r + 16384) >> 15).astype(np.uint8).reshape(256, -1)
    assert np.array_equal(oracle.gray(img), want)


@pytest.mark.parametrize("name", ["shift4", "shift23", "static", "cut", "noise", "static_noise", "flat"])
def test_farneback_pairs(golden, golden_arrays, oracle, name):
    c = next(p for p in golden["flow_pairs"] if p["name"] == name)
    fl = oracle.farneback(golden_arrays[f"pair_{name}.prev"], golden_arrays[f"pair_{name}.next"])
    m, v = oracle.mag_stats(fl)
    assert close(m, c["mean"], atol=flow_atol(c["mean"])), (m, c["mean"])
    assert close(v, c["var"], atol=flow_atol(c["mean"])), (v, c["var"])
    if f"pair_{name}.flow" in golden_arrays:
        ref = golden_arrays[f"pair_{name}.flow"]
        assert np.abs(fl - ref).max() <= 1e-3 * max(1.0, c["flow_abs_max"])
        # NumPy float32 pairwise statistics restated exactly
        assert oracle.mag_stats(ref) == (c["mean"], c["var"])


@pytest.mark.parametrize("name", ["clip_small", "clip_odd", "clip_nometa"])
def test_clip_end_to_end(golden, oracle, synth, name):
    c = golden["clips"][name]
    clip = synth.Clip(c["n_frames"], c["H"], c["W"], c["seed"], fps=c["fps"])
    step = oracle.sample_step(c["meta"].get("fps") or c["fps"])
    assert step == c["step"]
    frames = clip.sampled(step)
    assert sha(frames) == c["input_sha"]
    recs = oracle.frames_features(frames)
    assert [r["hash"].tolist() for r in recs] == c["hash_words"]
    assert [r["ham"] for r in recs[1:]] == c["ham"]
    assert [sha(r["thumb"]) for r in recs] == c["thumb_sha"]
    assert [r["lap_sum"] for r in recs] == c["lap_sum"] and [r["lap_sumsq"] for r in recs] == c["lap_sumsq"]
    for r, (m, v) in zip(recs[1:], c["flow_stats"]):
        assert close(r["flow_mean"], m, atol=flow_atol(m)) and close(r["flow_var"], v, atol=flow_atol(m)), (r["flow_mean"], m, r["flow_var"], v)
    duration = c["meta"].get("duration") or (c["n_frames"] / c["fps"])
    out = oracle.video_result(recs, c["H"], c["W"], c["fps"], duration)
    ref = c["result"]
    assert out["timeline"] is out["timeline_ai"]
    assert len(out["timeline"]) == len(ref["timeline"])
    for a, b in zip(out["timeline"], ref["timeline"]):
        assert close(a, b)
    for k, b in ref["summary"].items():
        a = out["summary"][k]
        if k in ("w", "h"):
            assert a == b and isinstance(a, int)
        elif k in ("dup_density", "scene_change_rate", "fps"):
            assert a == b
        else:
            assert close(a, b, atol=summary_flow_atol([m for m, _ in c["flow_stats"]], k)), (k, a, b)


def test_fuse_matches_reference(golden, oracle):
    import copy
    for c in golden["fuse"]:
        audio, video = copy.deepcopy(c["audio"]), copy.deepcopy(c["video"])
        assert oracle.hints(c["meta"]) == c["hints"]
        out = oracle.fuse(audio, video, c["hints"])
        assert out == c["out"]                       # float64 host arithmetic restated op for op: exact
        assert audio["timeline"] == c["audio_timeline_after"]   # in-place extension (fusion.py:20-21)
        assert video["timeline"] == c["video_timeline_after"]


def test_quirks(oracle):
    assert oracle.sample_step(25) == 12 and oracle.sample_step(30) == 15 and oracle.sample_step(60) == 30
    assert oracle.sample_step(0) == 15 and oracle.sample_step(1) == 1      # `fps or 30`; max(1, round(.5)=0)
    r = oracle.video_result([], 10, 10, 30.0, 2.5)                          # no samples: [0.5]*tlen, round(2.5)=2
    assert r["timeline"] == [0.5, 0.5] and r["summary"]["flow_mean"] == 0.0


def test_audio_oracle_reproduces_reference_dicts(synth):
    """oracle/audio_oracle.py against the dicts the reference's own audio.analyze returned (make_golden_audio.py)."""
    import hashlib
    import json
    import os
    from conftest import ROOT
    from oracle import audio_oracle
    with open(os.path.join(ROOT, "tests", "golden", "golden_audio.json")) as f:
        G = json.load(f)
    assert len(G["cases"]) == 10
    for c in G["cases"]:
        wav = synth.audio(c["seed"], c["n"], c["kind"])
        assert wav.dtype == np.float32 and hashlib.sha256(wav.tobytes()).hexdigest() == c["input_sha"], c["name"]
        with np.errstate(all="ignore"):
            assert audio_oracle.analyze_samples(wav, c["sr"]) == c["out"], c["name"]

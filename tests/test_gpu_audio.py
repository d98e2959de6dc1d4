"""K4 (avd_audio_windows): the audio analyzer's per-window features on the device against the oracle's NumPy restatement
of audio.py:41-61, and the whole analyzer (device features + host tail) against the dicts the reference's own analyze()
returned for the same seeded samples (tests/golden/golden_audio.json)."""
import json
import os

import numpy as np
import pytest

from conftest import ROOT, close

pytestmark = pytest.mark.gpu

# float64 features: the device spectrum is a direct DFT, the reference's a pocketfft rfft; both carry ~1e-13 of rounding
# relative to the largest bin.  North-star tolerance is 1e-4; the kernel is held to 1e-7.
FEAT_RTOL = 1e-7


@pytest.fixture(scope="module")
def eng():
    import torch
    assert torch.cuda.is_available(), "these tests need a GPU"
    from avd_b200.engine import Engine
    e = Engine(0)
    yield e
    e.close()


@pytest.fixture(scope="module")
def golden_audio():
    with open(os.path.join(ROOT, "tests", "golden", "golden_audio.json")) as f:
        return json.load(f)


@pytest.mark.parametrize("idx", range(10))
def test_window_features_vs_oracle_and_reference_dict(golden_audio, synth, eng, idx):
    from avd_b200 import audio
    from oracle import audio_oracle
    c = golden_audio["cases"][idx]
    wav = synth.audio(c["seed"], c["n"], c["kind"])
    with np.errstate(all="ignore"):
        want = np.stack(audio_oracle.window_features(wav, c["sr"]), axis=1)
    l0 = eng.launches
    got = audio.window_features(wav, c["sr"], engine=eng)
    assert eng.launches - l0 in (2, 3)                                 # spectrum (full windows / ragged window) + features, whatever the clip length
    assert got.shape == want.shape
    assert np.array_equal(got[:, 0], want[:, 0]), "rms (float32 arithmetic) must be exact"
    assert np.array_equal(got[:, 1], want[:, 1]), "zero-crossing rate must be exact"
    assert np.array_equal(got[:, 3], want[:, 3]), "roll-off bin must be exact"
    for col, name in ((2, "flatness"), (4, "centroid")):
        err = np.max(np.abs(got[:, col] - want[:, col]) / np.abs(want[:, col]))
        assert err <= FEAT_RTOL, (name, err)
    # whole analyzer against the reference's own output
    out, ref = audio.analyze_samples(wav, c["sr"], engine=eng), c["out"]
    assert out["scores"]["speech_ratio"] == ref["scores"]["speech_ratio"]
    assert close(out["scores"]["tts_like"], ref["scores"]["tts_like"])
    assert len(out["timeline"]) == len(ref["timeline"])
    assert all(close(a, b) for a, b in zip(out["timeline"], ref["timeline"]))
    for k, v in ref["flags_audio"].items():
        assert close(out["flags_audio"][k], v), k


def test_audio_edge_cases(eng):
    from avd_b200 import audio
    from avd_b200._lib import AvdError
    from oracle import audio_oracle
    assert audio.window_features(np.zeros(0, np.float32), 16000, engine=eng).shape == (0, 5)
    # a ragged window of ONE sample: hanning(1) = [1], one bin, zero-crossing rate NaN like np.mean of an empty diff
    wav = np.concatenate([np.linspace(-0.5, 0.5, 8000, dtype=np.float32), np.float32([0.25])])
    with np.errstate(all="ignore"):
        want = np.stack(audio_oracle.window_features(wav, 16000), axis=1)
    got = audio.window_features(wav, 16000, engine=eng)
    assert np.isnan(got[1, 1]) and np.isnan(want[1, 1])
    assert np.array_equal(got[:, 0], want[:, 0]) and np.array_equal(got[:, 3], want[:, 3])
    assert np.allclose(got[:, [2, 4]], want[:, [2, 4]], rtol=FEAT_RTOL, atol=0)
    # stereo input: first channel only (audio.py:34-35)
    st = np.stack([wav[:8000], -wav[:8000]], axis=1)
    assert audio.analyze_samples(st, 16000, engine=eng) == audio.analyze_samples(wav[:8000], 16000, engine=eng)
    with pytest.raises(AvdError):
        audio.window_features(np.zeros(48000, np.float32), 48000, engine=eng)   # 24000-sample windows: beyond the kernel's table
    with pytest.raises(AvdError):
        audio.window_features(np.zeros(70000, np.float32), 2, engine=eng)       # one-sample windows: more than 65535 of them
    # a window length that does not factor into two short DFTs (prime): every window takes the direct DFT
    w2 = (np.random.default_rng(3).integers(-9000, 9001, 3 * 4099 + 17).astype(np.float32) / np.float32(32768))
    with np.errstate(all="ignore"):
        want = np.stack(audio_oracle.window_features(w2, 8198), axis=1)                 # sr 8198 -> win 4099 (prime)
    got = audio.window_features(w2, 8198, engine=eng)
    assert got.shape == want.shape == (4, 5)
    assert np.array_equal(got[:, [0, 1, 3]], want[:, [0, 1, 3]]) and np.allclose(got[:, [2, 4]], want[:, [2, 4]], rtol=FEAT_RTOL, atol=0)
    eng.sync()


def test_audio_video_fuse_end_to_end(golden_audio, synth, eng):
    """Both analyzers on the device and the reference's fusion on top: the verdict equals the one computed from the
    reference's audio dict and the same video result (label, peaks; score within the rounding gate)."""
    import copy
    from avd_b200 import audio, fusion, hints as H, video
    c = golden_audio["cases"][0]                                        # 12 s of speech-like audio
    a_dev = audio.analyze_samples(synth.audio(c["seed"], c["n"], c["kind"]), c["sr"], engine=eng)
    clip = synth.Clip(360, 360, 640, seed=5, fps=30.0)
    meta = {"fps": 30.0, "width": 640, "height": 360, "duration": 12.0, "bit_rate": 800000}
    v = video.analyze_frames(clip.sampled(15), meta, engine=eng)
    hn = H.compute_hints(meta)
    got = fusion.fuse(a_dev, copy.deepcopy(v), hn, with_raw=True)
    want = fusion.fuse(copy.deepcopy(c["out"]), copy.deepcopy(v), hn, with_raw=True)
    assert got["result"]["label"] == want["result"]["label"] and got["peaks"] == want["peaks"]
    assert abs(got["_raw_score"] - want["_raw_score"]) <= 1e-3

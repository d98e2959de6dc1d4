"""CPU restatement of the reference's audio analyzer arithmetic -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows /root/reference/app/analyzers/audio.py:29-111 with the same NumPy calls in the same order, split at the seam
the device path uses: per-window features (audio.py:41-61) and the O(windows) tail (audio.py:63-111).  The decode
(audio.py:7-20: ffmpeg -> 16 kHz mono wav -> soundfile float32) is outside the path; inputs here are the float32 samples.
Pinned: tests/golden/golden_audio.json holds the dicts the reference's own analyze() returned for seeded inputs
(tests/golden/make_golden_audio.py feeds it the samples through its _extract_wav_16k hook); test_oracle_golden.py
requires this restatement to reproduce them exactly.
"""
import numpy as np


def window_features(wav: np.ndarray, sr: int):
    """audio.py:38-61 -> float64 arrays rms, zcr, flat, roll, sc (one entry per 0.5 s window, last one ragged)."""
    win = max(1, int(sr * 0.5)) if sr else 1
    rms, zcr, flat, roll, sc_cent = [], [], [], [], []
    for i in range(0, len(wav), win):
        seg = wav[i:i + win]
        if len(seg) == 0:
            continue
        rms.append(float(np.sqrt((seg ** 2).mean())))
        with np.errstate(all="ignore"):
            zc = np.mean(np.abs(np.diff(np.sign(seg)))) / 2.0 if len(seg) > 1 else np.float32(np.nan) / 2.0
        zcr.append(float(zc))
        winseg = seg * np.hanning(len(seg))
        spec = np.fft.rfft(winseg)
        mag = np.abs(spec) + 1e-9
        flat.append(float(np.exp(np.mean(np.log(mag))) / np.mean(mag)))
        cutoff = 0.85 * np.sum(mag)
        s, idx = 0.0, 0
        for k, m in enumerate(mag):
            s += m
            if s >= cutoff:
                idx = k
                break
        roll.append(float(idx) / max(1.0, len(mag)))
        freqs = np.linspace(0.0, 1.0, len(mag))
        sc_cent.append(float(np.sum(freqs * mag) / np.sum(mag)))
    return tuple(np.array(x, dtype=np.float64) for x in (rms, zcr, flat, roll, sc_cent))


def _norm01(x):
    x = np.asarray(x, dtype=float)
    if x.size == 0:
        return np.zeros(1)
    mn, mx = float(np.min(x)), float(np.max(x))
    return (x - mn) / (mx - mn + 1e-9)


def audio_result(feats, n_samples: int, sr: int) -> dict:
    """audio.py:36,63-111 from the per-window features."""
    rms, zcr, flat, roll, sc = feats
    dur = n_samples / sr if sr > 0 else 0.0
    rms_arr = rms if rms.size else np.zeros(1)
    zcr_arr = zcr if zcr.size else np.zeros(1)
    flat_arr = flat if flat.size else np.zeros(1)
    roll_arr = roll if roll.size else np.zeros(1)
    sc_arr = sc if sc.size else np.zeros(1)
    speech_thr = np.percentile(rms_arr, 60)
    speech_ratio = float(np.mean(rms_arr >= speech_thr))
    flat_mean, sc_var = float(np.mean(flat_arr)), float(np.var(sc_arr))
    roll_var, zcr_var = float(np.var(roll_arr)), float(np.var(zcr_arr))
    tts_base = 0.7 * flat_mean + 0.15 * (1.0 / (1e-6 + zcr_var)) + 0.15 * (1.0 / (1e-6 + roll_var))
    attenuation = 1.0 / (1.0 + 5.0 * (sc_var + roll_var + zcr_var))
    tts_like = float(np.clip(tts_base * attenuation, 0.0, 1.0))
    if sc_var + roll_var + zcr_var > 0.005:
        tts_like = float(min(tts_like, 0.90))
    dzcr = np.diff(np.concatenate([[zcr_arr[0]], zcr_arr]))
    droll = np.diff(np.concatenate([[roll_arr[0]], roll_arr]))
    tline = 0.5 * _norm01(flat_arr) + 0.3 * (1.0 - _norm01(dzcr ** 2)) + 0.2 * (1.0 - _norm01(np.abs(droll)))
    tline = np.clip(tline, 0.0, 1.0).tolist()
    tlen = int(max(1, round(dur)))
    if len(tline) < tlen:
        tline = tline + [tline[-1] if tline else 0.5] * (tlen - len(tline))
    else:
        tline = tline[:tlen]
    return {"scores": {"speech_ratio": speech_ratio, "tts_like": tts_like},
            "flags_audio": {"speech_ratio": speech_ratio, "tts_like": tts_like, "rms_var": float(np.var(rms_arr)),
                            "zcr_var": zcr_var, "roll_var": roll_var, "sc_var": sc_var},
            "timeline": tline}


def analyze_samples(wav: np.ndarray, sr: int) -> dict:
    wav = np.asarray(wav)
    if wav.ndim > 1:
        wav = wav[:, 0]
    return audio_result(window_features(wav, sr), len(wav), sr)

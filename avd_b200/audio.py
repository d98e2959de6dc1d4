"""Drop-in for app/analyzers/audio.py of the reference (same call, same return):

    analyze(path: str, meta: dict) -> {"scores": {...}, "flags_audio": {...}, "timeline": [...]}

The decode stays the reference's (audio.py:7-20: ffmpeg -> 16 kHz mono wav -> soundfile float32); when either is
missing the reference's own failure result comes back (audio.py:112-118).  The per-window features (audio.py:41-61:
RMS, zero crossings, Hann spectrum, flatness, roll-off, centroid) run on the B200 through libavd (K4,
`avd_audio_windows`); the O(windows) tail (audio.py:63-111) is float64 host arithmetic with the same NumPy calls.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import tempfile

import numpy as np

from .engine import lease


def window_features(wav: np.ndarray, sr: int, engine=None) -> np.ndarray:
    """(n_windows, 5) float64: rms, zcr, flat, roll, sc per 0.5 s window (audio.py:38-61), computed on the device."""
    wav = np.ascontiguousarray(wav, dtype=np.float32)
    win = max(1, int(sr * 0.5)) if sr else 1
    nwin = (len(wav) + win - 1) // win
    feats = np.zeros((nwin, 5), np.float64)
    if nwin:
        with lease(engine) as eng, eng._lock:
            eng._check(eng.L.avd_audio_windows(eng.h, C.c_void_p(wav.ctypes.data), len(wav), win,
                                               C.c_void_p(feats.ctypes.data), None), "avd_audio_windows")
    return feats


def _norm01(x):
    x = np.asarray(x, dtype=float)
    if x.size == 0:
        return np.zeros(1)
    mn, mx = float(np.min(x)), float(np.max(x))
    return (x - mn) / (mx - mn + 1e-9)


def result_from_features(feats: np.ndarray, n_samples: int, sr: int) -> dict:
    """Tail of audio.analyze (audio.py:36,63-111) from the per-window features."""
    feats = np.asarray(feats, dtype=np.float64).reshape(-1, 5)
    one = np.zeros(1)
    rms_arr, zcr_arr, flat_arr, roll_arr, sc_arr = (feats[:, i].copy() if len(feats) else one for i in range(5))
    dur = n_samples / sr if sr > 0 else 0.0
    speech_thr = np.percentile(rms_arr, 60)
    speech_ratio = float(np.mean(rms_arr >= speech_thr))
    flat_mean = float(np.mean(flat_arr))
    sc_var, roll_var, zcr_var = float(np.var(sc_arr)), float(np.var(roll_arr)), float(np.var(zcr_arr))
    tts_base = 0.7 * flat_mean + 0.15 * (1.0 / (1e-6 + zcr_var)) + 0.15 * (1.0 / (1e-6 + roll_var))
    attenuation = 1.0 / (1.0 + 5.0 * (sc_var + roll_var + zcr_var))
    tts_like = float(np.clip(tts_base * attenuation, 0.0, 1.0))
    if sc_var + roll_var + zcr_var > 0.005:          # cap when the variability is not negligible (audio.py:82-84)
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


def analyze_samples(wav: np.ndarray, sr: int, engine=None) -> dict:
    """Samples-level seam: ``wav`` is what the decode helper returns (float32, mono or (n, channels))."""
    wav = np.asarray(wav)
    if wav.ndim > 1:
        wav = wav[:, 0]                               # audio.py:34-35
    return result_from_features(window_features(wav, sr, engine), len(wav), sr)


def _extract_wav_16k(path: str):
    """audio.py:7-20, unchanged in behaviour: ffmpeg to a temporary 16 kHz mono wav, read as float32."""
    import soundfile as sf
    tmp = tempfile.NamedTemporaryFile(delete=False, suffix=".wav")
    tmp.close()
    proc = subprocess.run(["ffmpeg", "-y", "-i", path, "-ac", "1", "-ar", "16000", "-f", "wav", tmp.name],
                          stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    if proc.returncode != 0:
        os.unlink(tmp.name)
        raise RuntimeError("ffmpeg_convert_failed")
    try:
        wav, sr = sf.read(tmp.name, dtype="float32", always_2d=False)
    except Exception:
        raise RuntimeError("soundfile_read_failed")
    finally:
        try:
            os.unlink(tmp.name)
        except OSError:
            pass
    return wav, sr


def analyze(path: str, meta: dict, engine=None) -> dict:
    """Same contract as the reference's audio.analyze (audio.py:29-124), including its failure result."""
    try:
        wav, sr = _extract_wav_16k(path)
        return analyze_samples(wav, sr, engine)
    except Exception as e:                             # audio.py:112-118
        tlen = int(max(1, round(meta.get("duration") or 0.0)))
        return {"scores": {}, "flags_audio": {"error": str(e)}, "timeline": [0.5] * tlen}

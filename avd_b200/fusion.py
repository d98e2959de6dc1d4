"""Scoring tail: audio/video timelines + hints -> ai_score, label, confidence, peaks, binned timeline.
Mirrors fusion.fuse of the reference (fusion.py:16-109) -- same inputs, same output dict, same float64
operation order (so results are identical, not merely close), same in-place padding of the caller's
timeline lists (fusion.py:20-21).  O(seconds-of-video) scalars: stays on the host.
"""
from __future__ import annotations

import os
from dataclasses import dataclass

import numpy as np


@dataclass(frozen=True)
class Thresholds:
    real_max: float = float(os.getenv("THRESH_REAL_MAX", "0.35"))   # read once at import, fusion.py:4
    ai_min: float = float(os.getenv("THRESH_AI_MIN", "0.72"))       # fusion.py:5


THRESHOLDS = Thresholds()


def _pad(t: list, L: int) -> None:
    if len(t) < L:
        t += [t[-1] if t else 0.5] * (L - len(t))


def bin_timeline(ts) -> list:
    """3-tap moving average ("same"), clipped to [0,1] (fusion.py:7-14)."""
    if not ts:
        return []
    arr = np.array(ts, dtype=float)
    if len(arr) >= 3:
        arr = np.convolve(arr, np.ones(3) / 3.0, mode="same")
    return np.clip(arr, 0.0, 1.0).tolist()


def _weights(a, v, speech_ratio):
    w_a, w_v = 0.65, 0.25
    agree = 0.10 if np.sign(np.mean(a) - 0.5) == np.sign(np.mean(v) - 0.5) else 0.0
    if speech_ratio < 0.25:                       # little speech: lean on the video
        w_a *= 0.6
        w_v = max(0.2, 1.0 - w_a - agree)
    return w_a, w_v, agree


def _penalties(hints) -> float:
    p = 0.0
    if hints.get("compression", "normal") in ("heavy", "very_heavy"):
        p += 0.05
    if hints.get("bpp", 0.0) < 0.07:
        p += 0.05
    if hints.get("dup_avg", 0.0) > 0.2:
        p += 0.05
    return p


def _real_bonus(flow_mean, texture_var, sc_rate, dup_density) -> float:
    b = 0.0
    if flow_mean > 5.0 and texture_var > 200.0 and dup_density < 0.05:
        b -= 0.10
    if sc_rate > 0.7:
        b -= 0.05
    if sc_rate >= 0.9 and texture_var > 300.0 and dup_density < 0.02:
        b -= 0.08
    return b


def reason_for(label: str, dup_density: float, tts_like: float, hints: dict) -> str:
    """The explanation string of a label (fusion.py:82-100)."""
    comp = hints.get("compression", "normal")
    if label == "real":
        why = []
        if dup_density > 0.25:
            why.append("molti frame duplicati")
        if comp in ("heavy", "very_heavy"):
            why.append("compressione pesante")
        return "; ".join(why or ["segnali audio/video coerenti con ripresa reale"])
    if label == "ai":
        why = []
        if tts_like > 0.6:
            why.append("audio TTS-like elevato")
        if dup_density > 0.2:
            why.append("molti frame duplicati")
        if hints.get("video_has_signal", True) is False:
            why.append("segnali video deboli")
        return "; ".join(why or ["pattern e indizi coerenti con generazione AI"])
    return "segnali misti o neutri"


def fuse(audio: dict, video: dict, hints: dict, thresholds: Thresholds = THRESHOLDS, with_raw: bool = False) -> dict:
    a_t = audio.get("timeline") or []
    v_t = video.get("timeline") or video.get("timeline_ai") or []
    L = max(len(a_t), len(v_t), 1)
    _pad(a_t, L)
    _pad(v_t, L)
    a, v = np.array(a_t, dtype=float), np.array(v_t, dtype=float)

    flags = audio.get("flags_audio", {})
    speech_ratio = float(flags.get("speech_ratio", 0.0))
    tts_like = float(flags.get("tts_like", 0.0))
    w_a, w_v, agree = _weights(a, v, speech_ratio)
    pen = _penalties(hints)

    vs = video.get("summary", {}) or {}
    flow_mean = float(vs.get("flow_mean", 0.0))
    texture_var = float(vs.get("texture_var", 0.0))
    sc_rate = float(vs.get("scene_change_rate", 0.0))
    dup_density = float(vs.get("dup_density", 0.0))
    bonus = _real_bonus(flow_mean, texture_var, sc_rate, dup_density)
    if tts_like >= 0.95 and flow_mean > 8.0 and texture_var > 300.0 and dup_density < 0.05:
        w_a *= 0.55                                # TTS-like audio over clearly real footage
        w_v = max(0.25, 1.0 - w_a - agree)

    tl = np.clip((w_a * a + w_v * v + agree * (a + v) / 2.0) - pen + bonus, 0.0, 1.0)
    peaks = [i for i, x in enumerate(tl.tolist()) if x <= 0.25 or x >= 0.75]
    score = float(np.mean(tl))
    spread = float(np.std(tl))
    disagree = float(abs(np.mean(a) - np.mean(v)))
    conf = float(np.clip(0.20 + 2.2 * spread - pen - 0.5 * max(0.0, 0.3 - disagree), 0.10, 0.99))

    label = "real" if score <= thresholds.real_max else ("ai" if score >= thresholds.ai_min else "uncertain")
    reason = reason_for(label, dup_density, tts_like, hints)

    out = {"result": {"label": label, "ai_score": round(score, 2), "confidence": round(conf, 2), "reason": reason},
           "timeline_binned": bin_timeline(tl.tolist()), "peaks": peaks}
    if with_raw:
        out["_raw_score"] = score                  # pre-rounding score for the |d ai_score| <= 1e-3 gate
    return out


def neutral_audio(tlen: int) -> dict:
    """What the reference's audio analyzer returns when ffmpeg is unavailable (audio.py:112-118)."""
    return {"scores": {"speech_ratio": 0.0, "tts_like": 0.0}, "flags_audio": {"error": "unavailable"}, "timeline": [0.5] * int(tlen)}

"""NumPy/ctypes face of the CPU oracle.  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Two layers:

* per-primitive restatements (C, ``avd_oracle.c``) of the OpenCV/NumPy calls that
  ``/root/reference/app/analyzers/video.py`` makes -- each wrapper cites the call site;
* the scalar tail of the reference restated in plain Python/NumPy float64:
  per-sample score (video.py:54-57), clip summary (video.py:61-71), timeline (video.py:73-83),
  hints (heuristics_v2.py:1-18) and fusion (fusion.py:7-14,16-109).

Parity status: PINNED against outputs of the reference's own code executed in the build
container (cv2 4.13.0 / numpy 2.3.5, the reference pins 4.10.x / <2.0): tests/golden/*.npz, made by
tests/golden/make_golden.py.  The reference ships no tests or golden vectors of its own.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build() -> str:
    """Compile liboracle.so with gcc (idempotent)."""
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "avd_oracle.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        u8p, i64p, u32p, f32p = (C.POINTER(C.c_uint8), C.POINTER(C.c_int64), C.POINTER(C.c_uint32), C.POINTER(C.c_float))
        L.avd_o_gray.argtypes = [u8p, C.c_int, C.c_int, C.c_size_t, u8p]
        L.avd_o_area.argtypes = [u8p, C.c_int, C.c_int, C.c_int, u8p]; L.avd_o_area.restype = C.c_int
        L.avd_o_hash.argtypes = [u8p, u32p]
        L.avd_o_hamming.argtypes = [u32p, u32p]; L.avd_o_hamming.restype = C.c_int
        L.avd_o_thumb.argtypes = [u8p, C.c_int, C.c_int, C.c_int, u8p]
        L.avd_o_lapstats.argtypes = [u8p, C.c_int, C.c_int, i64p, i64p]
        L.avd_o_farneback.argtypes = [u8p, u8p, C.c_int, C.c_int, f32p]
        L.avd_o_magstats.argtypes = [f32p, C.c_size_t, f32p, f32p]
        _LIB = L
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def gray(bgr: np.ndarray) -> np.ndarray:
    """cv2.cvtColor(frame, COLOR_BGR2GRAY) -- video.py:5,43,51."""
    assert bgr.dtype == np.uint8 and bgr.ndim == 3 and bgr.shape[2] == 3
    bgr = np.ascontiguousarray(bgr)
    H, W, _ = bgr.shape
    out = np.empty((H, W), np.uint8)
    lib().avd_o_gray(_p(bgr, C.c_uint8), H, W, W * 3, _p(out, C.c_uint8))
    return out


def area32(g: np.ndarray, D: int = 32) -> np.ndarray:
    """cv2.resize(g, (32,32), interpolation=INTER_AREA) -- video.py:6."""
    g = np.ascontiguousarray(g)
    out = np.empty((D, D), np.uint8)
    if lib().avd_o_area(_p(g, C.c_uint8), g.shape[0], g.shape[1], D, _p(out, C.c_uint8)):
        raise ValueError("INTER_AREA oracle: source smaller than destination")
    return out


def hash_words(a32: np.ndarray) -> np.ndarray:
    """(g >= g.mean()) packed to 32 little-endian uint32 words -- video.py:7-8."""
    a32 = np.ascontiguousarray(a32)
    out = np.empty(32, np.uint32)
    lib().avd_o_hash(_p(a32, C.c_uint8), _p(out, C.c_uint32))
    return out


def hash_bits(words: np.ndarray) -> np.ndarray:
    """Unpack to the reference's 1024-vector of {0,1} uint8 (video.py:8 .flatten())."""
    return np.unpackbits(np.ascontiguousarray(words).view(np.uint8), bitorder="little")


def hamming(a: np.ndarray, b: np.ndarray) -> int:
    """int(np.sum(hsh ^ prev_hash)) -- video.py:38."""
    return int(lib().avd_o_hamming(_p(np.ascontiguousarray(a), C.c_uint32), _p(np.ascontiguousarray(b), C.c_uint32)))


def thumb320(g: np.ndarray, D: int = 320) -> np.ndarray:
    """cv2.resize(gray, (320,320)) default INTER_LINEAR -- video.py:43."""
    g = np.ascontiguousarray(g)
    out = np.empty((D, D), np.uint8)
    lib().avd_o_thumb(_p(g, C.c_uint8), g.shape[0], g.shape[1], D, _p(out, C.c_uint8))
    return out


def lap_sums(g: np.ndarray):
    """Integer sum and sum of squares of cv2.Laplacian(gray, CV_64F) -- video.py:52."""
    g = np.ascontiguousarray(g)
    s, ss = C.c_int64(), C.c_int64()
    lib().avd_o_lapstats(_p(g, C.c_uint8), g.shape[0], g.shape[1], C.byref(s), C.byref(ss))
    return int(s.value), int(ss.value)


def texture_from_sums(s: int, ss: int, npix: int) -> float:
    """float(Laplacian.var()) from the exact sums (population variance, float64)."""
    return (float(ss) - float(s) * float(s) / float(npix)) / float(npix)


def farneback(prev: np.ndarray, nxt: np.ndarray) -> np.ndarray:
    """cv2.calcOpticalFlowFarneback(prev, next, None, 0.5, 3, 15, 3, 5, 1.2, 0) -- video.py:45."""
    prev = np.ascontiguousarray(prev); nxt = np.ascontiguousarray(nxt)
    H, W = prev.shape
    out = np.empty((H, W, 2), np.float32)
    lib().avd_o_farneback(_p(prev, C.c_uint8), _p(nxt, C.c_uint8), H, W, _p(out, C.c_float))
    return out


def mag_stats(flow: np.ndarray):
    """float(np.mean(mag)), float(np.var(mag)) with mag = sqrt(fx^2+fy^2) in float32 -- video.py:46-48."""
    flow = np.ascontiguousarray(flow, np.float32)
    m, v = C.c_float(), C.c_float()
    lib().avd_o_magstats(_p(flow, C.c_float), flow.shape[0] * flow.shape[1], C.byref(m), C.byref(v))
    return float(m.value), float(v.value)


# ----------------------------------------------------------------------------------------------
# scalar tail
# ----------------------------------------------------------------------------------------------
def sample_step(fps) -> int:
    """video.py:19 -- Python round() is half-to-even."""
    return max(1, int(round((fps or 30) / 2)))


def ai_susp(tex: float, mot: float) -> float:
    """video.py:56-57."""
    return float(np.clip(1.0 - (tex / (tex + 1000.0)) * (1.0 + mot), 0.0, 1.0))


def frames_features(frames):
    """Per-sample integer/float features of a list of sampled BGR frames (video.py:36-52)."""
    recs = []
    prev_words, prev_thumb = None, None
    for f in frames:
        g = gray(f)
        words = hash_words(area32(g))
        th = thumb320(g)
        s, ss = lap_sums(g)
        r = {"hash": words, "thumb": th, "lap_sum": s, "lap_sumsq": ss,
             "ham": hamming(words, prev_words) if prev_words is not None else -1}
        if prev_thumb is not None:
            r["flow_mean"], r["flow_var"] = mag_stats(farneback(prev_thumb, th))
        prev_words, prev_thumb = words, th
        recs.append(r)
    return recs


def video_result(recs, H, W, fps, duration):
    """Tail of video.analyze (video.py:54-83) from per-sample records."""
    textures = [texture_from_sums(r["lap_sum"], r["lap_sumsq"], H * W) for r in recs]
    flow_means = [r["flow_mean"] for r in recs if "flow_mean" in r]
    flow_vars = [r["flow_var"] for r in recs if "flow_var" in r]
    dup = sum(1 for r in recs if r["ham"] == 0)
    total = len(recs)
    timeline = []
    k = 0
    for i, r in enumerate(recs):
        if "flow_mean" in r:
            k += 1
        mot = flow_means[k - 1] if k else 0.0
        timeline.append(ai_susp(textures[i], mot))
    summary = {
        "dup_density": float(dup / max(1, total - 1)),
        "scene_change_rate": float(np.mean(np.array(flow_vars) > 0.5)) if flow_vars else 0.0,
        "flow_mean": float(np.mean(flow_means)) if flow_means else 0.0,
        "flow_var": float(np.var(flow_means)) if flow_means else 0.0,
        "texture_var": float(np.var(textures)) if textures else 0.0,
        "w": int(W), "h": int(H), "fps": float(fps),
    }
    tlen = int(max(1, round(duration)))
    if len(timeline) < tlen:
        timeline = timeline + ([timeline[-1]] * (tlen - len(timeline)) if timeline else [0.5] * tlen)
    else:
        timeline = timeline[:tlen]
    return {"timeline": timeline, "summary": summary, "timeline_ai": timeline}


def hints(meta: dict) -> dict:
    """heuristics_v2.py:1-18."""
    w, h = meta.get("width") or 0, meta.get("height") or 0
    fps, br = meta.get("fps") or 0.0, meta.get("bit_rate") or 0
    pps = (w * h * fps) if w and h and fps else 0.0
    bpp = float(br) / max(1.0, pps)
    comp = "very_heavy" if bpp <= 0.04 else "heavy" if bpp <= 0.08 else "normal" if bpp <= 0.15 else "light"
    return {"w": w, "h": h, "fps": fps, "br": br, "bpp": round(bpp, 5), "compression": comp,
            "video_has_signal": (w * h) > 0 and fps > 0, "dup_avg": 0.0}


def fuse(audio: dict, video: dict, hnt: dict, thresh_real_max: float = 0.35, thresh_ai_min: float = 0.72, raw: bool = False):
    """fusion.py:16-109 (+ _bin_timeline :7-14).  Extends the input timelines in place like the reference."""
    a_t = audio.get("timeline") or []
    v_t = video.get("timeline") or video.get("timeline_ai") or []
    L = max(len(a_t), len(v_t), 1)
    if len(a_t) < L:
        a_t += [a_t[-1] if a_t else 0.5] * (L - len(a_t))
    if len(v_t) < L:
        v_t += [v_t[-1] if v_t else 0.5] * (L - len(v_t))
    a, v = np.array(a_t, dtype=float), np.array(v_t, dtype=float)
    w_a, w_v = 0.65, 0.25
    agree = 0.10 if np.sign(np.mean(a) - 0.5) == np.sign(np.mean(v) - 0.5) else 0.0
    flags = audio.get("flags_audio", {})
    speech, tts = float(flags.get("speech_ratio", 0.0)), float(flags.get("tts_like", 0.0))
    if speech < 0.25:
        w_a *= 0.6
        w_v = max(0.2, 1.0 - w_a - agree)
    comp, bpp, dup_h = hnt.get("compression", "normal"), hnt.get("bpp", 0.0), hnt.get("dup_avg", 0.0)
    pen = 0.0
    if comp in ("heavy", "very_heavy"): pen += 0.05
    if bpp < 0.07: pen += 0.05
    if dup_h > 0.2: pen += 0.05
    vs = video.get("summary", {}) or {}
    fm, tv = float(vs.get("flow_mean", 0.0)), float(vs.get("texture_var", 0.0))
    sc, dd = float(vs.get("scene_change_rate", 0.0)), float(vs.get("dup_density", 0.0))
    rb = 0.0
    if fm > 5.0 and tv > 200.0 and dd < 0.05: rb -= 0.10
    if sc > 0.7: rb -= 0.05
    if sc >= 0.9 and tv > 300.0 and dd < 0.02: rb -= 0.08
    if tts >= 0.95 and fm > 8.0 and tv > 300.0 and dd < 0.05:
        w_a *= 0.55
        w_v = max(0.25, 1.0 - w_a - agree)
    tl = np.clip((w_a * a + w_v * v + agree * (a + v) / 2.0) - pen + rb, 0.0, 1.0)
    peaks = [i for i, x in enumerate(tl.tolist()) if x <= 0.25 or x >= 0.75]
    score, spread = float(np.mean(tl)), float(np.std(tl))
    disagree = float(abs(np.mean(a) - np.mean(v)))
    conf = float(np.clip(0.20 + 2.2 * spread - pen - 0.5 * max(0.0, 0.3 - disagree), 0.10, 0.99))
    if score <= thresh_real_max:
        label = "real"
        why = (["molti frame duplicati"] if dd > 0.25 else []) + (["compressione pesante"] if comp in ("heavy", "very_heavy") else [])
        reason = "; ".join(why or ["segnali audio/video coerenti con ripresa reale"])
    elif score >= thresh_ai_min:
        label = "ai"
        why = (["audio TTS-like elevato"] if tts > 0.6 else []) + (["molti frame duplicati"] if dd > 0.2 else []) \
            + (["segnali video deboli"] if hnt.get("video_has_signal", True) is False else [])
        reason = "; ".join(why or ["pattern e indizi coerenti con generazione AI"])
    else:
        label, reason = "uncertain", "segnali misti o neutri"
    binned = tl.tolist()
    if binned:
        arr = np.array(binned, dtype=float)
        if len(arr) >= 3:
            arr = np.convolve(arr, np.ones(3) / 3.0, mode="same")
        binned = np.clip(arr, 0.0, 1.0).tolist()
    out = {"result": {"label": label, "ai_score": round(score, 2), "confidence": round(conf, 2), "reason": reason},
           "timeline_binned": binned, "peaks": peaks}
    if raw:
        out["_raw_score"] = score
    return out

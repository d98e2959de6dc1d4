"""Device-side scoring tail for batches of clips (K3, include/avd.h: avd_score_clips).

For a batch of independent clips (BASELINE configs[3]) the per-sample records stay in HBM after K1/K2 and one
small kernel produces, per clip, what video.py:54-81 and fusion.fuse (fusion.py:16-109) compute on the host in
the reference: the ai_susp timeline, the summary, the fused timeline, peaks, score, confidence, label.  Only
those compact results cross PCIe.  String work stays here: hints -> penalty, label -> reason.

    results = score_clip_batch(engine, clips, metas, audios=..., hints=...)     # frames in, verdicts out
    scores  = score_jobs(engine, jobs)                                           # lower level, explicit jobs
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import fusion
from ._lib import REC_DTYPE, AvdError

JOB_DTYPE = np.dtype([("rec_first", "<i4"), ("rec_count", "<i4"), ("npix", "<i4"), ("tlen", "<i4"), ("do_fuse", "<i4"),
                      ("audio_off", "<i4"), ("audio_len", "<i4"), ("video_given", "<i4"), ("vgiven_off", "<i4"),
                      ("vgiven_len", "<i4"), ("out_off", "<i4"), ("pad_", "<i4"),
                      ("speech_ratio", "<f8"), ("tts_like", "<f8"), ("pen", "<f8"),
                      ("thresh_real_max", "<f8"), ("thresh_ai_min", "<f8"),
                      ("g_flow_mean", "<f8"), ("g_texture_var", "<f8"), ("g_sc_rate", "<f8"), ("g_dup_density", "<f8")])
SCORE_DTYPE = np.dtype([("dup_density", "<f8"), ("scene_change_rate", "<f8"), ("flow_mean", "<f8"), ("flow_var", "<f8"),
                        ("texture_var", "<f8"), ("score", "<f8"), ("spread", "<f8"), ("conf", "<f8"),
                        ("ai_score", "<f8"), ("confidence", "<f8"),
                        ("label", "<i4"), ("vlen", "<i4"), ("L", "<i4"), ("n_peaks", "<i4")])
assert JOB_DTYPE.itemsize == 120 and SCORE_DTYPE.itemsize == 96
LABELS = ("real", "uncertain", "ai")


def timeline_length(duration) -> int:
    """video.py:73 -- Python round() (half to even) stays on the host."""
    return int(max(1, round(duration)))


class JobList:
    """Builds the job array and the concatenated input timelines of one avd_score_clips call."""

    def __init__(self, thresholds: fusion.Thresholds = fusion.THRESHOLDS):
        self.jobs, self.tl_in, self.n_out, self.thresholds = [], [], 0, thresholds

    def _put(self, values) -> tuple:
        off = len(self.tl_in)
        self.tl_in.extend(float(x) for x in values)
        return off, len(self.tl_in) - off

    def add(self, *, rec_first=0, rec_count=0, npix=0, tlen=1, video=None, audio=None, hints=None) -> int:
        """One clip: its records (rec_first, rec_count, npix, tlen) or a ready ``video`` dict; ``audio`` (dict as
        the reference's audio analyzer returns) switches the fuse part on."""
        j = np.zeros((), JOB_DTYPE)
        j["tlen"] = int(tlen)
        if video is not None:
            vs = video.get("summary", {}) or {}
            v_t = video.get("timeline") or video.get("timeline_ai") or []
            j["video_given"] = 1
            j["vgiven_off"], j["vgiven_len"] = self._put(v_t)
            j["g_flow_mean"] = float(vs.get("flow_mean", 0.0)); j["g_texture_var"] = float(vs.get("texture_var", 0.0))
            j["g_sc_rate"] = float(vs.get("scene_change_rate", 0.0)); j["g_dup_density"] = float(vs.get("dup_density", 0.0))
            vlen = int(j["vgiven_len"])
        else:
            j["rec_first"], j["rec_count"], j["npix"] = int(rec_first), int(rec_count), int(npix)
            vlen = int(tlen)
        L = vlen
        if audio is not None:
            flags = audio.get("flags_audio", {})
            j["do_fuse"] = 1
            j["audio_off"], j["audio_len"] = self._put(audio.get("timeline") or [])
            j["speech_ratio"] = float(flags.get("speech_ratio", 0.0)); j["tts_like"] = float(flags.get("tts_like", 0.0))
            j["pen"] = fusion._penalties(hints or {})
            j["thresh_real_max"], j["thresh_ai_min"] = self.thresholds.real_max, self.thresholds.ai_min
            L = max(int(j["audio_len"]), vlen, 1)
        j["out_off"] = self.n_out
        self.n_out += L
        self.jobs.append(j)
        return len(self.jobs) - 1

    def arrays(self):
        jobs = np.array(self.jobs, dtype=JOB_DTYPE) if self.jobs else np.zeros(0, JOB_DTYPE)
        return jobs, np.array(self.tl_in, dtype=np.float64), self.n_out


def score_jobs(engine, joblist: JobList, d_rec=None, n_rec: int = 0):
    """Runs K3.  d_rec: device pointer (int) of the records, or None for the records the engine's last
    ``analyze_host(..., keep_on_device=True)`` left on the device.
    Returns (jobs, scores[SCORE_DTYPE], vtimeline, fused, binned, peaks) as NumPy arrays; the per-position arrays are
    indexed from each job's out_off."""
    jobs, tl_in, n_out = joblist.arrays()
    n = len(jobs)
    scores = np.zeros(n, SCORE_DTYPE)
    vt, fu, bi = (np.zeros(max(n_out, 1), np.float64) for _ in range(3))
    pk = np.zeros(max(n_out, 1), np.uint8)
    if n == 0:
        return jobs, scores, vt, fu, bi, pk
    p = lambda a: C.c_void_p(a.ctypes.data)
    with engine._lock:
        engine._check(engine.L.avd_score_clips(engine.h, C.c_void_p(d_rec) if d_rec else None, int(n_rec), p(jobs), n,
                                               p(tl_in) if len(tl_in) else None, len(tl_in), p(scores), p(vt), p(fu), p(bi),
                                               p(pk), n_out, None), "avd_score_clips")
    return jobs, scores, vt, fu, bi, pk


def video_dict(job, score, vt, meta_w, meta_h, fps) -> dict:
    """What video.analyze returns (video.py:64-71,83), from one K3 score."""
    o, n = int(job["out_off"]), int(score["L"] if job["do_fuse"] else score["vlen"])
    timeline = vt[o:o + n].tolist()
    summary = {"dup_density": float(score["dup_density"]), "scene_change_rate": float(score["scene_change_rate"]),
               "flow_mean": float(score["flow_mean"]), "flow_var": float(score["flow_var"]),
               "texture_var": float(score["texture_var"]), "w": int(meta_w), "h": int(meta_h), "fps": float(fps)}
    return {"timeline": timeline, "summary": summary, "timeline_ai": timeline}


def fuse_dict(job, score, fu, bi, pk, hints: dict, with_raw: bool = False) -> dict:
    """What fusion.fuse returns (fusion.py:102-109), from one K3 score."""
    o, L = int(job["out_off"]), int(score["L"])
    label = LABELS[int(score["label"])]
    out = {"result": {"label": label, "ai_score": float(score["ai_score"]), "confidence": float(score["confidence"]),
                      "reason": fusion.reason_for(label, float(score["dup_density"]), float(job["tts_like"]), hints or {})},
           "timeline_binned": bi[o:o + L].tolist(), "peaks": np.flatnonzero(pk[o:o + L]).tolist()}
    if with_raw:
        out["_raw_score"] = float(score["score"])
    return out


def score_clip_batch(engine, clips, metas, audios=None, hints=None, with_raw: bool = False):
    """Frames in, verdicts out, for a batch of independent clips of one frame size: every sampled frame goes
    through K1/K2 in one host run whose records stay on the device, then K3 scores all clips in one launch.

    clips[i]: the sampled frames of clip i; metas[i]: its meta dict (fps, width, height, duration);
    audios[i] (optional): the audio analyzer's dict -- enables the fuse part, and its timeline list is padded in
    place exactly as fusion.fuse does (fusion.py:20); hints[i] (optional): compute_hints(meta).
    Returns a list of {"video": ..., "fusion": ... or None}.  After fusing, video["timeline"] has the padded
    length the reference leaves it with (fusion.py:21).
    """
    from .video import sample_step
    frames, jl, info = [], JobList(), []
    for ci, c in enumerate(clips):
        n = len(c)
        meta = metas[ci]
        fps = meta.get("fps") or 0.0
        H, W = (int(c[0].shape[0]), int(c[0].shape[1])) if n else (0, 0)
        duration = meta.get("duration") or (n * sample_step(fps) / fps if fps > 0 else 0.0)
        jl.add(rec_first=len(frames), rec_count=n, npix=H * W, tlen=timeline_length(duration),
               audio=None if audios is None else audios[ci], hints=None if hints is None else hints[ci])
        info.append((meta.get("width") or W, meta.get("height") or H, fps))
        frames.extend(c[i] for i in range(n))
    # the records of the host run stay in the context's own buffer: nobody else may run on this context between the
    # two calls (the engine's lock is re-entrant)
    with engine._lock:
        if frames:
            engine.analyze_host(frames, keep_on_device=True)
        jobs, scores, vt, fu, bi, pk = score_jobs(engine, jl, None, len(frames))
    out = []
    for ci in range(len(clips)):
        v = video_dict(jobs[ci], scores[ci], vt, *info[ci])
        f = None
        if audios is not None:
            f = fuse_dict(jobs[ci], scores[ci], fu, bi, pk, None if hints is None else hints[ci], with_raw)
            if audios[ci].get("timeline"):            # `audio.get("timeline") or []`: only a non-empty list is the caller's
                fusion._pad(audios[ci]["timeline"], int(scores[ci]["L"]))
        out.append({"video": v, "fusion": f})
    return out

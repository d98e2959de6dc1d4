"""Drop-in for app/analyzers/video.py of the reference (same call, same return):

    analyze(path: str, meta: dict) -> {"timeline": [...], "summary": {...}, "timeline_ai": <same list>}

Host decode stays what the reference uses (cv2.VideoCapture: grab every frame, retrieve the sampled ones,
video.py:11,27-34; for 8-bit 4:2:0 streams the same FFmpeg decoder is read before its BGR conversion, decode.py).  Everything computed per sampled frame runs on the B200 through libavd (engine.py);
the O(samples) scalar tail (video.py:54-83) is float64 host arithmetic restated op for op.
"""
from __future__ import annotations

import contextlib

import numpy as np

from . import decode
from ._lib import REC_DTYPE, REC_HAS_FLOW, REC_HAS_PREV
from .engine import lease

#: sampled frames per host->device run when decoding a file (bounds pinned memory; one-sample halo between runs)
DECODE_RUN = 48


def sample_step(fps) -> int:
    """video.py:19 -- ~2 analysed frames per second; Python round() is half-to-even (25 fps -> 12)."""
    return max(1, int(round((fps or 30) / 2)))


def _textures(recs: np.ndarray, npix: int) -> np.ndarray:
    """float(Laplacian.var()) per sample from the exact integer sums (video.py:52)."""
    s = recs["lap_sum"].astype(np.float64)
    ss = recs["lap_sumsq"].astype(np.float64)
    n = float(npix)
    return (ss - s * s / n) / n


def result_from_records(recs: np.ndarray, H: int, W: int, fps: float, duration: float, w=None, h=None) -> dict:
    """Tail of video.analyze (video.py:54-83) from the per-sample records of one clip, in sample order.

    recs[0] must be the clip's first sample (no predecessor); every later record carries flow statistics.
    ``w``/``h`` are the values echoed in the summary (the reference echoes meta or cap.get values, :15-16,70).
    """
    recs = np.asarray(recs).view(REC_DTYPE).reshape(-1)
    total = len(recs)
    tex = _textures(recs, H * W)
    has_flow = (recs["flags"] & REC_HAS_FLOW) != 0
    flow_means = recs["flow_mean"][has_flow].astype(np.float64)
    flow_vars = recs["flow_var"][has_flow].astype(np.float64)
    # mot = flow_means[-1] if flow_means else 0.0   (video.py:55): the latest flow mean seen so far
    if total:
        idx = np.cumsum(has_flow) - 1
        mot = np.where(idx >= 0, flow_means[np.maximum(idx, 0)] if len(flow_means) else 0.0, 0.0)
        susp = np.clip(1.0 - (tex / (tex + 1000.0)) * (1.0 + mot), 0.0, 1.0)
        timeline = [float(x) for x in susp]
    else:
        timeline = []
    dup = int(np.count_nonzero(((recs["flags"] & REC_HAS_PREV) != 0) & (recs["ham_prev"] == 0)))
    summary = {
        "dup_density": float(dup / max(1, total - 1)),
        "scene_change_rate": float(np.mean(flow_vars > 0.5)) if len(flow_vars) else 0.0,
        "flow_mean": float(np.mean(flow_means)) if len(flow_means) else 0.0,
        "flow_var": float(np.var(flow_means)) if len(flow_means) else 0.0,
        "texture_var": float(np.var(tex)) if total else 0.0,
        "w": int(W if w is None else w), "h": int(H if h is None else h), "fps": float(fps),
    }
    tlen = int(max(1, round(duration)))
    if len(timeline) < tlen:
        timeline = timeline + [timeline[-1]] * (tlen - len(timeline)) if timeline else [0.5] * tlen
    else:
        timeline = timeline[:tlen]
    return {"timeline": timeline, "summary": summary, "timeline_ai": timeline}


def records_from_frames(frames, engine=None) -> np.ndarray:
    """Per-sample records of already-sampled frames (host arrays), through the streaming host entry point."""
    with lease(engine) as eng:
        return eng.analyze_host(frames)


def analyze_frames(frames, meta: dict, engine=None) -> dict:
    """Frames-level seam: ``frames`` are the *sampled* frames (what video.py:32 retrieves), in order.

    meta: ``fps``, ``width``, ``height``, ``duration`` as in the reference (video.py:14-17); missing width /
    height default to the frame shape, missing duration to len(frames)*step/fps.
    """
    n = len(frames)
    fps = meta.get("fps") or 0.0
    if n == 0:
        return result_from_records(np.zeros(0, REC_DTYPE), 0, 0, fps, meta.get("duration") or 0.0,
                                   w=meta.get("width") or 0, h=meta.get("height") or 0)
    H, W = int(frames[0].shape[0]), int(frames[0].shape[1])
    duration = meta.get("duration") or (n * sample_step(fps) / fps if fps > 0 else 0.0)
    recs = records_from_frames(frames, engine)
    return result_from_records(recs, H, W, fps, duration, w=meta.get("width") or W, h=meta.get("height") or H)


class _RunPipeline:
    """Decode of run k+1 overlaps the device work of run k (the reference's loop is strictly serial, video.py:27-58).

    Two pinned stages of DECODE_RUN + 1 frame slots (slot 0 = the one-sample halo: the last frame of the previous run,
    which seeds the hash / thumbnail chain, video.py:20,26,41,49).  The decode loop fills one stage while a worker thread
    has the other one inside `Engine.analyze_host` (ctypes releases the GIL; so do cv2 and libav while decoding).
    """

    def __init__(self, eng, stages, frame_nbytes: int, fmt: str, shape):
        from concurrent.futures import ThreadPoolExecutor
        self.eng, self.fmt, self.shape, self.nb = eng, fmt, shape, frame_nbytes
        self.stages = [st[:(DECODE_RUN + 1) * frame_nbytes].reshape(DECODE_RUN + 1, frame_nbytes) for st in stages]
        self.cur, self.fill, self.have_halo = 0, 0, False
        self.pending = [None, None]                          # future of the run that last used stage i
        self.runs = []
        self.pool = ThreadPoolExecutor(max_workers=1, thread_name_prefix="avd-run")

    def slot(self) -> np.ndarray:
        """The staging slot the next sampled frame is decoded into."""
        return self.stages[self.cur][1 + self.fill]

    def pushed(self):
        self.fill += 1
        if self.fill == DECODE_RUN:
            self.flush()

    def _frames(self, stage, lo, hi):
        if self.fmt == "bgr":
            H, W = self.shape
            return stage[lo:hi].reshape(hi - lo, H, W, 3)
        return stage[lo:hi]

    def _run(self, stage, lo, hi, drop_first):
        recs = self.eng.analyze_host(self._frames(stage, lo, hi), fmt=self.fmt, shape=self.shape)
        return recs[1:] if drop_first else recs

    def flush(self):
        if self.fill == 0:
            return
        stage, other = self.stages[self.cur], self.stages[1 - self.cur]
        lo = 0 if self.have_halo else 1
        fut = self.pool.submit(self._run, stage, lo, 1 + self.fill, self.have_halo)
        self.pending[self.cur] = fut
        self.runs.append(fut)
        # the other stage is about to be refilled: its previous run must be done with it, then it gets this run's last
        # frame as its halo slot
        if self.pending[1 - self.cur] is not None:
            self.pending[1 - self.cur].result()
            self.pending[1 - self.cur] = None
        np.copyto(other[0], stage[self.fill])
        self.cur, self.fill, self.have_halo = 1 - self.cur, 0, True

    def finish(self) -> np.ndarray:
        self.flush()
        parts = [f.result() for f in self.runs]
        return np.concatenate(parts) if parts else np.zeros(0, REC_DTYPE)

    def close(self):
        """Waits for the worker: the stages must not go back to the staging cache while a run still reads them."""
        self.pool.shutdown(wait=True)


def analyze(path: str, meta: dict, engine=None) -> dict:
    """Same contract as the reference's video.analyze (video.py:10-83).

    Decode is the reference's own (cv2.VideoCapture: grab every frame, retrieve the sampled ones).  For 8-bit 4:2:0
    streams the sampled pictures are taken from the decoder *before* the BGR conversion `retrieve()` does on the host
    and converted on the device instead, bit for bit (decode.LibavSource, csrc/k0_yuv.cu): half the bytes cross PCIe.
    """
    import cv2  # host decode only: the reference's own decoder
    cap = cv2.VideoCapture(path)
    if not cap.isOpened():
        return {"timeline": [], "summary": {}, "timeline_ai": []}
    fps = meta.get("fps") or cap.get(cv2.CAP_PROP_FPS) or 0.0
    w = meta.get("width") or int(cap.get(cv2.CAP_PROP_FRAME_WIDTH) or 0)
    h = meta.get("height") or int(cap.get(cv2.CAP_PROP_FRAME_HEIGHT) or 0)
    duration = meta.get("duration") or (cap.get(cv2.CAP_PROP_FRAME_COUNT) / fps if fps > 0 else 0.0)
    step = sample_step(fps)

    yuv = decode.open_yuv_source(path, cap)
    if yuv is not None:
        try:
            first = yuv.grab()                               # learns the pixel format; may refuse the stream
        except decode.DecodeUnavailable:
            yuv.release()
            yuv = None
    try:
        with lease(engine) as eng:
            if yuv is not None:
                cap.release()
                recs, shape = _decode_loop_yuv(eng, yuv, first, step)
            else:
                recs, shape = _decode_loop_bgr(eng, cap, step)
    finally:
        cap.release()
        if yuv is not None:
            yuv.release()
    H, W = shape if shape else (int(h), int(w))
    return result_from_records(recs, H, W, fps, duration, w=w, h=h)


def _decode_loop_bgr(eng, cap, step):
    pipe, shape, index = None, None, 0
    with contextlib.ExitStack() as stack:
        try:
            while True:
                if not cap.grab():
                    break
                if index % step == 0:
                    ok, frame = cap.retrieve()
                    if not ok:
                        break                                   # video.py:33: keep what we have
                    if pipe is None:
                        shape = frame.shape
                        nb = int(frame.size)
                        stages = [stack.enter_context(eng.staging((DECODE_RUN + 1) * nb)) for _ in range(2)]
                        pipe = _RunPipeline(eng, stages, nb, "bgr", (shape[0], shape[1]))
                    if frame.shape != shape:
                        raise ValueError("frame size changed mid-stream")
                    np.copyto(pipe.slot().reshape(shape), frame)
                    pipe.pushed()
                index += 1
            recs = pipe.finish() if pipe else np.zeros(0, REC_DTYPE)
        finally:
            if pipe:
                pipe.close()
    return recs, (shape[0], shape[1]) if shape else None


def _decode_loop_yuv(eng, src, have_first: bool, step):
    pipe, shape, index = None, None, 0
    with contextlib.ExitStack() as stack:
        try:
            ok = have_first
            while ok:
                if index % step == 0:
                    if pipe is None:
                        shape = (src.height, src.width)
                        nb = src.frame_nbytes()
                        stages = [stack.enter_context(eng.staging((DECODE_RUN + 1) * nb)) for _ in range(2)]
                        pipe = _RunPipeline(eng, stages, nb, src.fmt, shape)
                    if not src.retrieve_into(pipe.slot()):
                        break
                    pipe.pushed()
                index += 1
                ok = src.grab()
            recs = pipe.finish() if pipe else np.zeros(0, REC_DTYPE)
        finally:
            if pipe:
                pipe.close()
    return recs, shape


def analyze_frames_batch(clips, metas, engine=None):
    """A batch of independent clips (BASELINE configs[3]): ``clips[i]`` are the sampled frames of clip i (all
    clips one frame size), ``metas[i]`` its meta dict.  All frames go through the device in ONE run; the
    records at clip boundaries are then detached from their (foreign) predecessor, which is exactly the state
    video.py starts every clip with (prev_hash = None, prev_frame_small = None; video.py:20,26).
    Returns one result dict per clip."""
    frames, starts = [], []
    for c in clips:
        starts.append(len(frames))
        frames.extend(c[i] for i in range(len(c)))
    recs = records_from_frames(frames, engine) if frames else np.zeros(0, REC_DTYPE)
    out = []
    for ci, c in enumerate(clips):
        n = len(c)
        r = recs[starts[ci]:starts[ci] + n].copy()
        if n:
            r["ham_prev"][0] = -1
            r["flags"][0] = 0
            r["flow_mean"][0] = 0.0
            r["flow_var"][0] = 0.0
        meta = metas[ci]
        fps = meta.get("fps") or 0.0
        H, W = (int(c[0].shape[0]), int(c[0].shape[1])) if n else (0, 0)
        duration = meta.get("duration") or (n * sample_step(fps) / fps if fps > 0 else 0.0)
        out.append(result_from_records(r, H, W, fps, duration, w=meta.get("width") or W, h=meta.get("height") or H))
    return out

"""Service-side wiring of the drop-ins: what the reference's request handler does around the two
analyzers (`api.py:143-169`), with the two things the GPU path makes worth doing:

* `analyze_path` runs the audio and the video analyzer of one request side by side (the reference awaits them one after
  the other, `api.py:145-146`): their decodes (ffmpeg, cv2) overlap; each analyzer call leases its own context from the device's engine pool
  (engine.lease), so concurrent requests overlap on the GPU as well.
  Failures are replaced by the reference's neutral results and recorded in the hints exactly as `_safe_audio` /
  `_safe_video` do (`api.py:118-141`).
* `RequestBatcher` coalesces requests that arrive together on the server's pool threads (`api.py:133`) into one
  `video.analyze_frames_batch` call: one host->device run and one K2 pass for all of them instead of one per request.
  Results are the ones each request would have got alone (clips are independent: `video.py:20,26`).

Nothing here touches HTTP, uploads, probing or URLs: that stays the reference's `api.py`.
"""
from __future__ import annotations

import threading
import time
from concurrent.futures import ThreadPoolExecutor

from . import audio as audio_an
from . import fusion as fusion_an
from . import hints as hx
from . import video as video_an


def _tlen(meta: dict) -> int:
    return int(max(1, round(meta.get("duration") or 0.0)))


def safe_audio(path: str, meta: dict, engine=None):
    """`_safe_audio` of the reference (`api.py:118-128`) without the asyncio plumbing."""
    try:
        return audio_an.analyze(path, meta, engine=engine), {}
    except Exception as e:                                   # noqa: BLE001 - the reference catches everything here
        name = e.__class__.__name__
        return {"scores": {}, "flags_audio": {"error": name}, "timeline": [0.5] * _tlen(meta)}, {"audio_error": name}


def safe_video(path: str, meta: dict, engine=None):
    """`_safe_video` of the reference (`api.py:130-141`)."""
    try:
        return video_an.analyze(path, meta, engine=engine), {}
    except Exception as e:                                   # noqa: BLE001
        name = e.__class__.__name__
        t = [0.5] * _tlen(meta)
        return {"timeline": t, "summary": {"error": name}, "timeline_ai": list(t)}, {"video_error": name}


_pool = ThreadPoolExecutor(max_workers=8, thread_name_prefix="avd-analyzer")


def analyze_path(path: str, meta: dict, engine=None) -> dict:
    """The analysis part of `_analyze_path` (`api.py:143-160`): hints, both analyzers (concurrently), fuse.
    Returns the same keys the reference puts in its response for them."""
    hints = hx.compute_hints(meta, path)
    fa = _pool.submit(safe_audio, path, meta, engine)
    video, v_hint = safe_video(path, meta, engine)
    audio, a_hint = fa.result()
    hints.update(a_hint)
    hints.update(v_hint)
    fused = fusion_an.fuse(audio, video, hints)
    return {"hints": hints, "video": video, "audio": audio, "result": fused["result"],
            "timeline_binned": fused["timeline_binned"], "peaks": fused["peaks"]}


class RequestBatcher:
    """Coalesces concurrent `analyze_frames(frames, meta)` calls into `video.analyze_frames_batch` calls.

    The first caller to arrive becomes the leader of a batch: it waits up to `max_wait_ms` for followers (or until
    `max_frames` sampled frames are queued), runs the whole batch through the engine once and hands every caller its own
    result.  Frames of one batch must share a shape (the engine streams one shape per run); a request with another shape
    starts its own batch.  A failure of the batch is raised in every caller of that batch.
    """

    class _Req:
        __slots__ = ("frames", "meta", "done", "result", "error")

        def __init__(self, frames, meta):
            self.frames, self.meta = frames, meta
            self.done = threading.Event()
            self.result = self.error = None

    def __init__(self, engine=None, max_wait_ms: float = 0.5, max_frames: int = 512, batch_fn=None):
        self.engine, self.max_wait, self.max_frames = engine, max_wait_ms / 1e3, int(max_frames)
        self._batch_fn = batch_fn or video_an.analyze_frames_batch
        self._lock = threading.Lock()
        self._pending = {}                    # shape -> list of requests waiting for a leader to take them
        self.batches = []                     # sizes of the batches run so far (for tests / metrics)

    def analyze_frames(self, frames, meta: dict) -> dict:
        if len(frames) == 0:
            return video_an.analyze_frames(frames, meta, engine=self.engine)
        shape = tuple(frames[0].shape)
        req = self._Req(frames, meta)
        with self._lock:
            queue = self._pending.setdefault(shape, [])
            leader = not queue
            queue.append(req)
        if leader:
            deadline = time.monotonic() + self.max_wait
            while time.monotonic() < deadline:
                with self._lock:
                    if sum(len(r.frames) for r in self._pending[shape]) >= self.max_frames:
                        break
                time.sleep(self.max_wait / 8)
            with self._lock:
                batch = self._pending.pop(shape)
            try:
                results = self._batch_fn([r.frames for r in batch], [r.meta for r in batch], engine=self.engine)
                for r, res in zip(batch, results):
                    r.result = res
            except Exception as e:                           # noqa: BLE001 - handed to every caller of the batch
                for r in batch:
                    r.error = e
            self.batches.append(len(batch))
            for r in batch:
                r.done.set()
        req.done.wait()
        if req.error is not None:
            raise req.error
        return req.result

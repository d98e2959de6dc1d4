"""CPU baseline body -- TEST/BENCH INFRASTRUCTURE, NOT PRODUCT CODE.

The reference's per-sample loop body (video.py:36-57) restated around the SAME library calls it makes
(cv2.cvtColor / resize / calcOpticalFlowFarneback / Laplacian, np.mean / np.var), operating on frames that
are already in memory.  /root/reference does not exist on the GPU box, so bench.py's `cpu_baseline` and
`--impl reference` legs time this port (kind "port") with OpenCV's own SIMD kernels doing the arithmetic --
i.e. the reference's real CPU cost, not a scalar C loop.  Used only by bench.py and tests.
"""
from __future__ import annotations

import numpy as np


class SampleChain:
    """Carried state of the sampled-frame loop: previous hash and previous 320x320 gray (video.py:20,26)."""

    def __init__(self, collect_ints: bool = False):
        self.prev_hash = None
        self.prev_small = None
        self.dup = 0
        self.flow_means, self.flow_vars, self.textures, self.timeline = [], [], [], []
        # parity passes only (never the timed baseline): the integer intermediates behind the floats above
        self.collect_ints = collect_ints
        self.hash_words, self.hams, self.lap_sums = [], [], []

    def push(self, frame: np.ndarray, emit: bool = True):
        import cv2
        g32 = cv2.resize(cv2.cvtColor(frame, cv2.COLOR_BGR2GRAY), (32, 32), interpolation=cv2.INTER_AREA)   # :5-6
        hsh = (g32 >= g32.mean()).astype(np.uint8).flatten()                                                # :7-8
        if emit and self.prev_hash is not None and int(np.sum(hsh ^ self.prev_hash)) == 0:                 # :37-40
            self.dup += 1
        if emit and self.collect_ints:
            self.hash_words.append(np.packbits(hsh, bitorder="little").view(np.uint32).copy())
            self.hams.append(int(np.sum(hsh ^ self.prev_hash)) if self.prev_hash is not None else -1)
        self.prev_hash = hsh
        small = cv2.resize(cv2.cvtColor(frame, cv2.COLOR_BGR2GRAY), (320, 320))                             # :43
        if emit and self.prev_small is not None:
            flow = cv2.calcOpticalFlowFarneback(self.prev_small, small, None, 0.5, 3, 15, 3, 5, 1.2, 0)     # :45
            mag = np.sqrt(flow[..., 0] ** 2 + flow[..., 1] ** 2)                                            # :46
            self.flow_means.append(float(np.mean(mag)))
            self.flow_vars.append(float(np.var(mag)))
        self.prev_small = small
        if not emit:
            return
        lap = cv2.Laplacian(cv2.cvtColor(frame, cv2.COLOR_BGR2GRAY), cv2.CV_64F)                            # :51
        tex = float(lap.var())                                                                              # :52
        self.textures.append(tex)
        if self.collect_ints:
            self.lap_sums.append((int(lap.sum()), int((lap * lap).sum())))
        mot = self.flow_means[-1] if self.flow_means else 0.0
        self.timeline.append(float(np.clip(1.0 - (tex / (tex + 1000.0)) * (1.0 + mot), 0.0, 1.0)))          # :56-57


def run_chunk(frames, halo=None, collect_ints: bool = False):
    """Body over a contiguous chunk of sampled frames (optionally seeded by a halo frame).  Returns
    (textures, flow_means, flow_vars, dup, timeline); with collect_ints (parity passes, untimed) additionally
    (hash words per sample, Hamming to the predecessor or -1, (sum, sum of squares) of the Laplacian)."""
    import cv2
    cv2.setNumThreads(1)
    ch = SampleChain(collect_ints)
    if halo is not None:
        ch.push(halo, emit=False)
    for f in frames:
        ch.push(f)
    out = (ch.textures, ch.flow_means, ch.flow_vars, ch.dup, ch.timeline)
    return out + (ch.hash_words, ch.hams, ch.lap_sums) if collect_ints else out


_SHARED = {}


def _worker(args):
    lo, hi = args
    fr = _SHARED["frames"]
    return run_chunk(fr[lo:hi], fr[lo - 1] if lo > 0 else None, _SHARED.get("collect_ints", False))


def run_parallel(frames, procs: int):
    """All host cores: one process per contiguous chunk with a one-sample halo (fork; frames shared COW)."""
    import multiprocessing as mp
    n = len(frames)
    procs = max(1, min(procs, n))
    base, extra = divmod(n, procs)
    bounds, lo = [], 0
    for r in range(procs):
        hi = lo + base + (1 if r < extra else 0)
        bounds.append((lo, hi)); lo = hi
    _SHARED["frames"] = frames
    if procs == 1:
        parts = [_worker(bounds[0])]
    else:
        with mp.get_context("fork").Pool(procs) as pool:
            parts = pool.map(_worker, bounds)
    tex = [t for p in parts for t in p[0]]
    fm = [t for p in parts for t in p[1]]
    fv = [t for p in parts for t in p[2]]
    return tex, fm, fv, sum(p[3] for p in parts), [t for p in parts for t in p[4]]


class Pool:
    """Persistent fork pool so repeated timed steps do not pay process start-up."""

    def __init__(self, frames, procs: int):
        import multiprocessing as mp
        self.n = len(frames)
        self.procs = max(1, min(procs, self.n // 2 or 1))   # >= 2 samples per chunk: halo overhead <= 1 frame in 3
        base, extra = divmod(self.n, self.procs)
        self.bounds, lo = [], 0
        for r in range(self.procs):
            hi = lo + base + (1 if r < extra else 0)
            self.bounds.append((lo, hi)); lo = hi
        _SHARED["frames"] = frames
        self.pool = mp.get_context("fork").Pool(self.procs) if self.procs > 1 else None

    def step(self):
        parts = self.pool.map(_worker, self.bounds) if self.pool else [_worker(self.bounds[0])]
        self.last = parts                       # per-chunk outputs of this pass (bench.py's parity check reads them)
        return sum(len(p[0]) for p in parts)

    def outputs(self):
        """The last pass, concatenated in sample order: dict of textures, flow_means, flow_vars, dup, timeline.  Every chunk
        after the first was seeded by its halo sample, so flow lists have one entry per sample except the very first."""
        parts = self.last
        cat = lambda k: [x for p in parts for x in p[k]]
        return {"textures": cat(0), "flow_means": cat(1), "flow_vars": cat(2), "dup": sum(p[3] for p in parts), "timeline": cat(4)}

    def close(self):
        if self.pool:
            self.pool.close(); self.pool.join()

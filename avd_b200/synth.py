"""Seeded synthetic clips -- NumPy only, integer arithmetic only, so the same seed gives
the same bytes on every machine (no cv2, no float reductions).

A clip is a texture panned at ``pan`` px/frame (default 1: 15 px between 2 Hz samples, which Farneback can
track; the 4 px/frame of SURVEY s8d saturates ai_susp at 0) with per-frame +-3 noise, scripted so that every branch of
the analyzer is exercised (reference behaviour cited from /root/reference/app/analyzers/video.py):

    [0.00,0.30)  texture A panning                      -> ordinary flow / texture
    [0.30,0.40)  frozen frame (exact duplicates)        -> aHash Hamming == 0, dup counter (video.py:37-40)
    [0.40,0.60)  scene cut to texture B, panning        -> flow_var > 0.5 at the cut (video.py:62)
    [0.60,0.75)  flat, low-texture frames               -> Laplacian var -> 0, ai_susp -> 1 (video.py:56)
    [0.75,0.90)  static texture B with fresh noise      -> near-zero flow
    [0.90,1.00)  texture A panning again                -> second cut
"""
from __future__ import annotations

import numpy as np

_BINOM9 = np.array([1, 8, 28, 56, 70, 56, 28, 8, 1], np.uint32)  # /256, sigma ~ 1.41; applied twice ~ sigma 2


def _blur_axis(a: np.ndarray, axis: int) -> np.ndarray:
    """9-tap binomial blur along ``axis`` with edge replication, exact in uint32."""
    pad = [(0, 0)] * a.ndim
    pad[axis] = (4, 4)
    p = np.pad(a, pad, mode="edge").astype(np.uint32)
    n = a.shape[axis]
    acc = np.zeros(a.shape, np.uint32)
    for k, w in enumerate(_BINOM9):
        sl = [slice(None)] * a.ndim
        sl[axis] = slice(k, k + n)
        acc += w * p[tuple(sl)]
    return ((acc + 128) >> 8).astype(np.uint8)


def texture(H: int, W: int, seed: int, tag: int) -> np.ndarray:
    """Blurred uniform noise, contrast-stretched back to a wide range. H x W x 3 uint8."""
    rng = np.random.default_rng([seed, tag, H, W])
    t = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    for _ in range(2):
        t = _blur_axis(_blur_axis(t, 0), 1)
    # integer contrast stretch x4 about 128 (blurring shrinks the range), clipped
    gain = 4 if tag == 0 else 7  # texture B is harsher so texture_var of a clip is not degenerate
    return np.clip((t.astype(np.int32) - 128) * gain + 128, 0, 255).astype(np.uint8)


class Clip:
    """Random-access synthetic clip: ``clip[i]`` is frame i (H x W x 3 uint8 BGR)."""

    def __init__(self, n_frames: int, H: int, W: int, seed: int, fps: float = 30.0, pan: int = 1,
                 noise: int = 3, events: bool = True, margin: int = 256):
        self.n, self.H, self.W, self.seed, self.fps = int(n_frames), int(H), int(W), int(seed), float(fps)
        self.pan, self.noise, self.events, self.margin = int(pan), int(noise), bool(events), int(margin)
        self._tex = {}

    @property
    def duration(self) -> float:
        return self.n / self.fps

    def _texture(self, tag: int) -> np.ndarray:
        if tag not in self._tex:
            self._tex[tag] = texture(self.H, self.W + self.margin, self.seed, tag)
        return self._tex[tag]

    def _offset(self, i: int) -> int:
        """Triangle-wave pan inside the margin (continuous motion, never wraps)."""
        period = 2 * self.margin
        p = (i * self.pan) % period
        return p if p <= self.margin else period - p

    def _noisy(self, base: np.ndarray, i: int) -> np.ndarray:
        if not self.noise:
            return np.ascontiguousarray(base)
        rng = np.random.default_rng([self.seed, 7, i])
        nz = rng.integers(-self.noise, self.noise + 1, base.shape, dtype=np.int16)
        return np.clip(base.astype(np.int16) + nz, 0, 255).astype(np.uint8)

    def __len__(self):
        return self.n

    def __getitem__(self, i: int) -> np.ndarray:
        if not 0 <= i < self.n:
            raise IndexError(i)
        H, W = self.H, self.W
        t = i / self.n if self.events else 0.0
        if t < 0.30 or t >= 0.90:
            o = self._offset(i)
            return self._noisy(self._texture(0)[:, o:o + W], i)
        if t < 0.40:  # frozen: every frame equals the first frame of the segment, bit for bit
            j = int(np.ceil(0.30 * self.n))
            o = self._offset(j)
            return self._noisy(self._texture(0)[:, o:o + W], j)
        if t < 0.60:
            o = self._offset(i)
            return self._noisy(self._texture(1)[:, o:o + W], i)
        if t < 0.75:  # flat: slow horizontal ramp, no noise
            ramp = (np.arange(W, dtype=np.int32) * 16 // max(1, W - 1) + 120 + (i % 2)).astype(np.uint8)
            return np.ascontiguousarray(np.broadcast_to(ramp[None, :, None], (H, W, 3)))
        o = self._offset(int(np.ceil(0.75 * self.n)))  # static + fresh noise
        return self._noisy(self._texture(1)[:, o:o + W], i)

    def sampled(self, step: int) -> np.ndarray:
        """Frames with index % step == 0 (what video.py:31 analyses), stacked n x H x W x 3."""
        idx = range(0, self.n, step)
        out = np.empty((len(idx), self.H, self.W, 3), np.uint8)
        for k, i in enumerate(idx):
            out[k] = self[i]
        return out


def noise_frames(n: int, H: int, W: int, seed: int) -> np.ndarray:
    """Worst case for the per-frame kernel: i.i.d. uniform bytes (max Laplacian energy, incompressible)."""
    rng = np.random.default_rng([seed, 99, H, W])
    return rng.integers(0, 256, (n, H, W, 3), dtype=np.uint8)


def bgr_to_i420(bgr: np.ndarray) -> np.ndarray:
    """Synthetic *decoder output*: a dense planar 4:2:0 picture (H*W luma bytes, then the (H/2)x(W/2) U and V planes) made
    from a BGR frame with the usual BT.601 limited-range integer matrix and a 2x2 chroma average.  This plays the
    encoder + decoder for tests and benchmarks; it is not part of the analysed path -- whatever bytes it yields simply ARE
    the decoded picture whose conversion to BGR (cv2.VideoCapture.retrieve, video.py:32) the device reproduces."""
    b, g, r = (bgr[..., i].astype(np.int32) for i in range(3))
    y = ((66 * r + 129 * g + 25 * b + 128) >> 8) + 16
    u = ((-38 * r - 74 * g + 112 * b + 128) >> 8) + 128
    v = ((112 * r - 94 * g - 18 * b + 128) >> 8) + 128
    H, W = y.shape
    sub = lambda p: ((p[0::2, 0::2] + p[0::2, 1::2] + p[1::2, 0::2] + p[1::2, 1::2] + 2) >> 2)
    out = np.empty(H * W * 3 // 2, np.uint8)
    out[:H * W] = np.clip(y, 0, 255).reshape(-1)
    out[H * W:H * W + H * W // 4] = np.clip(sub(u), 0, 255).reshape(-1)
    out[H * W + H * W // 4:] = np.clip(sub(v), 0, 255).reshape(-1)
    return out


def audio(seed: int, n_samples: int, kind: str = "speech") -> np.ndarray:
    """Mono 16 kHz test audio as float32 in [-1, 1) (what soundfile hands audio.py:15 of the reference), built from
    int16 PCM with integer arithmetic only, so every machine regenerates the same samples.

    kinds: "noise" white noise; "speech" noise bursts of varying level separated by exact silence (voiced / unvoiced /
    pause pattern, 0.2-0.9 s segments); "tone" a 125-sample triangle wave with a little noise; "silence" zeros;
    "quiet" noise of +-2 LSB.
    """
    rng = np.random.default_rng([seed, 411])
    n = int(n_samples)
    if kind == "silence":
        pcm = np.zeros(n, np.int32)
    elif kind == "noise":
        pcm = rng.integers(-12000, 12001, n, dtype=np.int32)
    elif kind == "quiet":
        pcm = rng.integers(-2, 3, n, dtype=np.int32)
    elif kind == "tone":
        t = np.arange(n, dtype=np.int32) % 125
        tri = np.where(t < 63, t * 300, (125 - t) * 300) - 9300
        pcm = tri + rng.integers(-40, 41, n, dtype=np.int32)
    elif kind == "speech":
        pcm = np.zeros(n, np.int32)
        pos, state = 0, 0
        while pos < n:
            seg = int(rng.integers(3200, 14401))
            end = min(n, pos + seg)
            if state % 3 == 0:      # voiced: low-passed noise (running mean of 8) at a random level
                amp = int(rng.integers(4000, 16001))
                x = rng.integers(-amp, amp + 1, end - pos + 8, dtype=np.int32)
                c = np.cumsum(x, dtype=np.int64)
                pcm[pos:end] = ((c[8:] - c[:-8]) // 4).astype(np.int32)[:end - pos]
            elif state % 3 == 1:    # unvoiced: white noise
                amp = int(rng.integers(500, 3001))
                pcm[pos:end] = rng.integers(-amp, amp + 1, end - pos, dtype=np.int32)
            pos, state = end, state + 1
    else:
        raise ValueError(kind)
    return (np.clip(pcm, -32768, 32767).astype(np.float32) / np.float32(32768.0)).astype(np.float32)

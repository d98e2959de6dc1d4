"""Engine: one libavd context on one GPU.  Thin, typed wrapper over the C ABI (include/avd.h).

Device memory and streams come from PyTorch (plumbing only): device-resident entry points take uint8 CUDA
tensors and run on torch's current stream, so torch.cuda.Event timing and torch.distributed collectives see
them.  The host entry point (`analyze_host`) is torch-free: NumPy frames in, NumPy records out.
"""
from __future__ import annotations

import ctypes as C
import json
import threading

import numpy as np

from . import _lib
from ._lib import REC_DTYPE, THUMB, AvdError


class Engine:
    def __init__(self, device: int = 0):
        self.L = _lib.load()
        if self.L.avd_device_count() <= 0:
            raise AvdError("no CUDA device visible: avd-b200 has no CPU fallback")
        h = C.c_void_p()
        rc = self.L.avd_create(int(device), C.byref(h))
        if rc != 0:
            raise AvdError(f"avd_create(device={device}) failed with status {rc} (needs an sm_100 GPU)")
        self.h, self.device = h, int(device)
        self._lock = threading.Lock()   # a context is single-threaded (include/avd.h)
        self._graphs = {}

    # -- plumbing ------------------------------------------------------------------------------------
    graph_launches = 0           # kernel launches replayed through CUDA graphs (not seen by avd_launch_count)

    def close(self):
        if self._graphs:
            self._graphs.clear()
        if getattr(self, "h", None):
            self.L.avd_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int, what: str):
        if rc != 0:
            raise AvdError(f"{what}: status {rc}: {self.L.avd_last_error(self.h).decode(errors='replace')}")

    def sync(self):
        self._check(self.L.avd_sync(self.h), "avd_sync")

    @property
    def launches(self) -> int:
        return int(self.L.avd_launch_count(self.h))

    def set_k1_mode(self, mode: int):
        self._check(self.L.avd_set_k1_mode(self.h, int(mode)), "avd_set_k1_mode")

    def profile(self, on: bool):
        self._check(self.L.avd_profile_enable(self.h, int(bool(on))), "avd_profile_enable")

    def profile_read(self) -> dict:
        buf = C.create_string_buffer(1 << 16)
        self._check(self.L.avd_profile_read(self.h, buf, len(buf)), "avd_profile_read")
        return json.loads(buf.value.decode())

    def pinned(self, shape, dtype=np.uint8) -> np.ndarray:
        """Page-locked host array (cudaHostAlloc) for full-rate cudaMemcpyAsync; freed with the engine."""
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = self.L.avd_host_alloc(self.h, n)
        if not p:
            raise AvdError("avd_host_alloc failed")
        buf = (C.c_uint8 * n).from_address(p)
        return np.frombuffer(buf, dtype=dtype).reshape(shape)

    @staticmethod
    def _stream():
        import torch
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)

    # -- device-resident path ---------------------------------------------------------------------------
    def frames_features(self, frames, prev_hash=None, want_area32: bool = False):
        """K1 on an (n,H,W,3) uint8 CUDA tensor.  Returns (records uint8[n,160], thumbs uint8[n,320,320][, area32])."""
        import torch
        assert frames.is_cuda and frames.dtype == torch.uint8 and frames.dim() == 4 and frames.shape[3] == 3
        assert frames.stride(3) == 1 and frames.stride(2) == 3, "pixels must be packed BGR"
        n, H, W, _ = frames.shape
        rec = torch.empty((n, REC_DTYPE.itemsize), dtype=torch.uint8, device=frames.device)
        thumbs = torch.empty((n, THUMB, THUMB), dtype=torch.uint8, device=frames.device)
        area = torch.empty((n, 32, 32), dtype=torch.uint8, device=frames.device) if want_area32 else None
        ph = C.c_void_p(prev_hash.data_ptr()) if prev_hash is not None else None
        with self._lock:
            self._check(self.L.avd_frames_features(self.h, C.c_void_p(frames.data_ptr()), n, H, W, frames.stride(1),
                                                   frames.stride(0) if n > 1 else H * frames.stride(1), ph,
                                                   C.c_void_p(rec.data_ptr()), C.c_void_p(thumbs.data_ptr()),
                                                   C.c_void_p(area.data_ptr()) if want_area32 else None, self._stream()),
                        "avd_frames_features")
        return (rec, thumbs, area) if want_area32 else (rec, thumbs)

    def flow_stats(self, thumbs, rec, prev_thumb=None):
        """K2: fills rec[i].flow_mean/var for consecutive thumbnails (and (prev_thumb, thumbs[0]) if given)."""
        with self._lock:
            self._check(self.L.avd_flow_stats(self.h, C.c_void_p(thumbs.data_ptr()), thumbs.shape[0],
                                              C.c_void_p(prev_thumb.data_ptr()) if prev_thumb is not None else None,
                                              C.c_void_p(rec.data_ptr()), self._stream()), "avd_flow_stats")

    def flow_fields(self, prev, nxt):
        """Dense flow (n,320,320,2) float32 for explicit pairs of (n,320,320) uint8 CUDA tensors."""
        import torch
        n = prev.shape[0]
        out = torch.empty((n, THUMB, THUMB, 2), dtype=torch.float32, device=prev.device)
        with self._lock:
            self._check(self.L.avd_flow_fields(self.h, C.c_void_p(prev.data_ptr()), C.c_void_p(nxt.data_ptr()), n,
                                               C.c_void_p(out.data_ptr()), self._stream()), "avd_flow_fields")
        return out

    def gray(self, frame):
        import torch
        H, W, _ = frame.shape
        out = torch.empty((H, W), dtype=torch.uint8, device=frame.device)
        with self._lock:
            self._check(self.L.avd_gray(self.h, C.c_void_p(frame.data_ptr()), H, W, frame.stride(0),
                                        C.c_void_p(out.data_ptr()), self._stream()), "avd_gray")
        return out

    def analyze_device(self, frames, prev_frame=None):
        """Whole device path for frames resident in HBM: K1 + K2.  Returns NumPy records (n,) REC_DTYPE.
        prev_frame: optional (H,W,3) CUDA tensor, the sample before frames[0] (chunk halo)."""
        import torch
        if prev_frame is not None:
            rec0, th0 = self.frames_features(prev_frame[None])
            prev_hash = rec0.view(torch.int32)[0, 6:38]          # hash words: bytes 24..152
            rec, thumbs = self.frames_features(frames, prev_hash=prev_hash)
            self.flow_stats(thumbs, rec, prev_thumb=th0[0])
        else:
            rec, thumbs = self.frames_features(frames)
            self.flow_stats(thumbs, rec)
        return rec

    def analyze_device_graph(self, frames, prev_frame=None):
        """analyze_device through a CUDA graph: the ~40 kernel launches of the device path are captured once per
        (buffers, shape) and replayed, which removes most of the launch gaps between the small Farneback kernels.
        The inputs must stay at the same addresses between calls (a resident batch, a reused staging buffer); the
        returned record tensor is owned by the graph and overwritten by the next call."""
        import torch
        key = (frames.data_ptr(), tuple(frames.shape), tuple(frames.stride()),
               None if prev_frame is None else prev_frame.data_ptr())
        g = self._graphs.get(key)
        if g is None:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                self.analyze_device(frames, prev_frame)      # warm-up: tables, workspaces, function attributes
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            l0 = self.launches
            with torch.cuda.graph(graph):
                rec = self.analyze_device(frames, prev_frame)
            g = self._graphs[key] = (graph, rec, self.launches - l0)
        g[0].replay()
        self.graph_launches += g[2]
        return g[1]

    @staticmethod
    def records_to_numpy(rec) -> np.ndarray:
        return rec.cpu().numpy().view(REC_DTYPE).reshape(-1)

    # -- host path (the public, torch-free call) --------------------------------------------------------
    def analyze_host(self, frames, keep_on_device: bool = False):
        """frames: sequence of (H,W,3) uint8 arrays (same shape, rows contiguous) or one (n,H,W,3) array,
        ideally in pinned memory (`Engine.pinned`).  Streams them through K1/K2 and returns records.
        keep_on_device: leave the records in HBM for `scoring.score_jobs` (K3) and return None."""
        n = len(frames)
        out = None if keep_on_device else np.zeros(n, REC_DTYPE)
        if n == 0:
            return out
        f0 = frames[0]
        H, W = int(f0.shape[0]), int(f0.shape[1])
        pitch = int(f0.strides[0])
        ptrs = (C.c_void_p * n)()
        for i in range(n):
            f = frames[i]
            if f.dtype != np.uint8 or f.shape != (H, W, 3) or f.strides[2] != 1 or f.strides[1] != 3 or f.strides[0] != pitch:
                raise AvdError("analyze_host: frames must be uint8 (H,W,3), packed BGR, one common shape and row pitch")
            ptrs[i] = f.ctypes.data
        with self._lock:
            self._check(self.L.avd_analyze_host(self.h, ptrs, n, H, W, pitch,
                                                C.c_void_p(out.ctypes.data) if out is not None else None), "avd_analyze_host")
        return out


_default = {}
_default_lock = threading.Lock()


def default_engine(device: int = 0) -> Engine:
    """Process-wide engine per device (created lazily)."""
    with _default_lock:
        if device not in _default:
            _default[device] = Engine(device)
        return _default[device]

"""Engine: one libavd context on one GPU.  Thin, typed wrapper over the C ABI (include/avd.h).

Device memory and streams come from PyTorch (plumbing only): device-resident entry points take uint8 CUDA
tensors and run on torch's current stream, so torch.cuda.Event timing and torch.distributed collectives see
them.  The host entry point (`analyze_host`) is torch-free: NumPy frames in, NumPy records out.
"""
from __future__ import annotations

import contextlib
import ctypes as C
import json
import os
import threading
from collections import OrderedDict

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
        self._lock = threading.RLock()  # a context is single-threaded (include/avd.h); re-entrant so that a caller can hold
        #                                 it across several entry points that share context state (scoring.score_clip_batch)
        self._graphs = OrderedDict()    # CUDA graphs of the device path, LRU, keyed by buffers + workspace generation
        self._staging = {}              # nbytes -> free pinned staging buffers (reused across calls, freed with the context)

    # -- plumbing ------------------------------------------------------------------------------------
    graph_launches = 0           # kernel launches replayed through CUDA graphs (not seen by avd_launch_count)

    GRAPH_CACHE = 8              # cached graphs per engine
    STAGING_CACHE_BYTES = 4 << 30   # pinned bytes kept for reuse per engine; beyond that, returned buffers are freed

    def close(self):
        if getattr(self, "_graphs", None):
            self._graphs.clear()
        if getattr(self, "_staging", None):
            self._staging.clear()         # the regions themselves are released by avd_destroy
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

    @property
    def ws_generation(self) -> int:
        """Bumped by the library whenever one of its scratch buffers moves (include/avd.h: avd_ws_generation)."""
        return int(self.L.avd_ws_generation(self.h))

    def check_guards(self) -> int:
        """Corrupted guard bytes around the library's device workspaces (0 = intact); raises on a CUDA error."""
        with self._lock:
            n = int(self.L.avd_check_guards(self.h))
        if n < 0:
            self._check(n, "avd_check_guards")
        return n

    def set_k1_mode(self, mode: int):
        self._check(self.L.avd_set_k1_mode(self.h, int(mode)), "avd_set_k1_mode")

    def profile(self, on: bool):
        self._check(self.L.avd_profile_enable(self.h, int(bool(on))), "avd_profile_enable")

    def profile_read(self) -> dict:
        buf = C.create_string_buffer(1 << 16)
        self._check(self.L.avd_profile_read(self.h, buf, len(buf)), "avd_profile_read")
        return json.loads(buf.value.decode())

    def pinned(self, shape, dtype=np.uint8) -> np.ndarray:
        """Page-locked host array (cudaHostAlloc) for full-rate cudaMemcpyAsync.  The region lives until
        `free_pinned(array)` or until the engine is closed (avd_destroy releases what is left) -- long-lived callers
        that need a scratch buffer per call should use `staging()` instead, which reuses regions."""
        n = max(1, int(np.prod(shape)) * np.dtype(dtype).itemsize)
        with self._lock:
            p = self.L.avd_host_alloc(self.h, n)
        if not p:
            raise AvdError(f"avd_host_alloc({n} bytes) failed")
        buf = (C.c_uint8 * n).from_address(p)
        return np.frombuffer(buf, dtype=dtype)[:int(np.prod(shape))].reshape(shape)

    def free_pinned(self, arr: np.ndarray):
        with self._lock:
            self.L.avd_host_free(self.h, C.c_void_p(arr.ctypes.data))

    @property
    def pinned_regions(self) -> int:
        """Live cudaHostAlloc regions of this context (leak checks)."""
        return int(self.L.avd_host_alloc_count(self.h))

    @contextlib.contextmanager
    def staging(self, nbytes: int):
        """A pinned uint8 scratch buffer of `nbytes` for the duration of the `with` block.  Buffers are cached per
        size and handed out again by later calls, so a server calling `video.analyze` per request allocates pinned memory
        once per frame size, not once per request."""
        nbytes = int(nbytes)
        with self._lock:
            free = self._staging.get(nbytes)
            buf = free.pop() if free else None
        if buf is None:
            buf = self.pinned((nbytes,))
        try:
            yield buf
        finally:
            with self._lock:
                cached = sum(k * len(v) for k, v in self._staging.items())
                if self.h and cached + nbytes <= self.STAGING_CACHE_BYTES:
                    self._staging.setdefault(nbytes, []).append(buf)
                elif self.h:
                    self.L.avd_host_free(self.h, C.c_void_p(buf.ctypes.data))

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

    def yuv420_to_bgr(self, yuv, H: int, W: int, full_range: bool = False, pitch: int = 0):
        """K0 on (n, 1.5*H*W) uint8 CUDA tensor of dense I420 frames -> (n, H, pitch) uint8 packed BGR rows."""
        import torch
        n = yuv.shape[0]
        pitch = int(pitch) or 3 * W
        out = torch.zeros((n, H, pitch), dtype=torch.uint8, device=yuv.device)
        with self._lock:
            self._check(self.L.avd_yuv420_to_bgr(self.h, C.c_void_p(yuv.data_ptr()), yuv.stride(0), n, H, W, int(bool(full_range)),
                                                 C.c_void_p(out.data_ptr()), pitch, H * pitch, self._stream()), "avd_yuv420_to_bgr")
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
        """Whole device path for frames resident in HBM: K1 + K2.  Returns the (n, 160) uint8 CUDA tensor of records.
        prev_frame: optional (H,W,3) CUDA tensor, the sample before frames[0] (chunk halo).  When the halo frame sits
        directly in front of `frames` in memory (one staging buffer: [halo, f0, f1, ...]) the n + 1 frames go through ONE
        K1 launch and one K2 pass and the halo's record is dropped; otherwise the halo gets its own one-frame K1 launch."""
        import torch
        if prev_frame is not None:
            n, fb = frames.shape[0], frames.stride(0)
            if (n > 0 and prev_frame.is_contiguous() and frames[0].is_contiguous() and prev_frame.shape == frames.shape[1:]
                    and prev_frame.data_ptr() + fb == frames.data_ptr() and fb == prev_frame.numel()):
                both = torch.as_strided(prev_frame, (n + 1,) + tuple(frames.shape[1:]), frames.stride())
                rec, thumbs = self.frames_features(both)
                self.flow_stats(thumbs, rec)
                return rec[1:]
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
        with self._lock:
            key = (frames.data_ptr(), tuple(frames.shape), tuple(frames.stride()),
                   None if prev_frame is None else prev_frame.data_ptr())
            g = self._graphs.get(key)
            # a graph holds the addresses of the library's grow-only scratch buffers: when any of them has moved since the
            # capture (a later call with more frames, another shape, analyze_host on this engine ...) it must not be replayed
            if g is not None and g[3] != self.ws_generation:
                del self._graphs[key]
                g = None
            if g is None:
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    self.analyze_device(frames, prev_frame)      # warm-up: tables, workspaces, function attributes
                torch.cuda.current_stream().wait_stream(side)
                torch.cuda.synchronize()
                graph = torch.cuda.CUDAGraph()
                l0, gen = self.launches, self.ws_generation
                with torch.cuda.graph(graph):
                    rec = self.analyze_device(frames, prev_frame)
                if self.ws_generation != gen:
                    raise AvdError("workspace moved during graph capture")     # cannot happen after the warm-up run
                g = self._graphs[key] = (graph, rec, self.launches - l0, gen)
                while len(self._graphs) > self.GRAPH_CACHE:
                    self._graphs.popitem(last=False)
            else:
                self._graphs.move_to_end(key)
            g[0].replay()
            self.graph_launches += g[2]
            return g[1]

    @staticmethod
    def records_to_numpy(rec) -> np.ndarray:
        return rec.cpu().numpy().view(REC_DTYPE).reshape(-1)

    # -- host path (the public, torch-free call) --------------------------------------------------------
    def analyze_host(self, frames, keep_on_device: bool = False, fmt: str = "bgr", shape=None):
        """frames: sequence of (H,W,3) uint8 arrays (same shape, rows contiguous), one (n,H,W,3) array, or a prepared
        `HostBatch`; ideally in pinned memory (`Engine.pinned`).  Streams them through K1/K2 and returns records.
        keep_on_device: leave the records in HBM for `scoring.score_jobs` (K3) and return None.
        fmt "i420" / "j420" (with shape=(H, W)): every frame is a dense planar 4:2:0 picture of 1.5*H*W bytes as the
        decoder emits it (limited / full range); the colour conversion of cv2's retrieve() then runs on the device."""
        b = frames if isinstance(frames, HostBatch) else HostBatch(frames, fmt, shape)
        out = None if keep_on_device else np.zeros(b.n, REC_DTYPE)
        if b.n == 0:
            return out
        o = C.c_void_p(out.ctypes.data) if out is not None else None
        with self._lock:
            if b.fmt == "bgr":
                self._check(self.L.avd_analyze_host(self.h, b.ptrs, b.n, b.H, b.W, b.pitch, o), "avd_analyze_host")
            else:
                self._check(self.L.avd_analyze_host_yuv420(self.h, b.ptrs, b.n, b.H, b.W, int(b.fmt == "j420"), o), "avd_analyze_host_yuv420")
        return out


class HostBatch:
    """Validated frame pointers of one `analyze_host` run.  Building it costs a few microseconds per frame (Python), so a
    caller that re-uses its staging buffers (a server, the decode pipeline of video.analyze) prepares it once per buffer:
    it stays valid while the frames keep their addresses and shapes (the contents may change)."""

    def __init__(self, frames, fmt: str = "bgr", shape=None):
        if fmt not in ("bgr", "i420", "j420"):
            raise AvdError("frame format must be 'bgr', 'i420' or 'j420'")
        self.fmt, self.n, self.keepalive = fmt, len(frames), frames
        self.H = self.W = self.pitch = 0
        self.ptrs = None
        if self.n == 0:
            return
        if fmt == "bgr":
            f0 = frames[0]
            self.H, self.W, self.pitch = int(f0.shape[0]), int(f0.shape[1]), int(f0.strides[0])
            per_frame = (self.H, self.W, 3)
            ok = lambda f: (f.dtype == np.uint8 and f.shape == per_frame and f.strides[2] == 1 and f.strides[1] == 3
                            and f.strides[0] == self.pitch)
            err = "analyze_host: frames must be uint8 (H,W,3), packed BGR, one common shape and row pitch"
        else:
            if shape is None:
                raise AvdError("analyze_host: 4:2:0 frames need shape=(H, W)")
            self.H, self.W = int(shape[0]), int(shape[1])
            self.pitch = 3 * self.W
            nb = self.H * self.W * 3 // 2
            ok = lambda f: f.dtype == np.uint8 and f.size == nb and f.flags["C_CONTIGUOUS"]
            err = "analyze_host: 4:2:0 frames must be contiguous uint8 arrays of 1.5*H*W bytes"
        if isinstance(frames, np.ndarray) and frames.ndim >= 2:
            # one array of n frames: validate one frame, compute the addresses arithmetically
            if not ok(frames[0]):
                raise AvdError(err)
            addr = frames.ctypes.data + np.arange(self.n, dtype=np.uint64) * np.uint64(frames.strides[0])
        else:
            addr = np.empty(self.n, np.uint64)
            for i in range(self.n):
                f = frames[i]
                if not ok(f):
                    raise AvdError(err)
                addr[i] = f.__array_interface__["data"][0]
        self._addr = addr
        self.ptrs = addr.ctypes.data_as(C.POINTER(C.c_void_p))


# ---- process-wide engines ------------------------------------------------------------------------------------
# The reference runs its analyzer on pool threads (api.py:133), so several requests can be inside video.analyze at once.
# A context is single-threaded; instead of serialising every request on one context, each device has a small pool of
# contexts (own streams, own workspaces) and a request leases one for its duration: concurrent small requests then
# overlap on the GPU (the Farneback levels of a short clip are latency-bound and leave most SMs idle).
POOL_SIZE = max(1, int(os.environ.get("AVD_ENGINES_PER_DEVICE", "4")))


class EnginePool:
    def __init__(self, device: int = 0, size: int = POOL_SIZE):
        self.device, self.size = int(device), int(size)
        self._cv = threading.Condition()
        self._free, self._all = [], []

    @contextlib.contextmanager
    def lease(self):
        with self._cv:
            while not self._free and len(self._all) >= self.size:
                self._cv.wait()
            if self._free:
                eng = self._free.pop()
            else:
                eng = Engine(self.device)
                self._all.append(eng)
        try:
            yield eng
        finally:
            with self._cv:
                self._free.append(eng)
                self._cv.notify()

    def first(self) -> Engine:
        with self._cv:
            if not self._all:
                eng = Engine(self.device)
                self._all.append(eng)
                self._free.append(eng)
            return self._all[0]

    def close(self):
        with self._cv:
            for e in self._all:
                e.close()
            self._all, self._free = [], []


_pools = {}
_pools_lock = threading.Lock()


def engine_pool(device: int = 0) -> EnginePool:
    with _pools_lock:
        if device not in _pools:
            _pools[device] = EnginePool(device)
        return _pools[device]


def default_engine(device: int = 0) -> Engine:
    """The first engine of the device's pool (created lazily).  Calls on it are serialised by its lock; request handlers
    should use `lease()`."""
    return engine_pool(device).first()


@contextlib.contextmanager
def lease(engine=None, device: int = 0):
    """`with lease(engine) as eng:` -- the given engine, or one leased from the device's pool for the block."""
    if engine is not None:
        yield engine
    else:
        with engine_pool(device).lease() as eng:
            yield eng

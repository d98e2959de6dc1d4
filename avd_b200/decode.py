"""Host decode for the drop-in `video.analyze(path, meta)` -- the reference's grab/retrieve loop (video.py:11,27-34).

`video.analyze` has two frame sources: `cv2.VideoCapture`, exactly what the reference uses (`retrieve()` returns packed BGR,
3 bytes / pixel), and, defined here:

* `LibavSource` -- the same FFmpeg libraries OpenCV's backend decodes with (the copies bundled in the opencv wheel, or the
  system's), driven directly through their C API with ctypes, so that the decoder's native planar 4:2:0 picture can be
  handed on *before* the colour conversion `retrieve()` would do on the host.  The GPU then performs that conversion
  bit for bit (csrc/k0_yuv.cu), and only 1.5 bytes per pixel cross PCIe (first step).  Where PyAV is
  installed the same planes are available from `av.VideoFrame.planes`; this image has no PyAV, hence the ctypes shim.

`LibavSource` is used only when it is provably equivalent: 8-bit yuv420p / yuvj420p, even width and height, no rotation
metadata (OpenCV auto-rotates), libavformat major >= 60 (struct offsets below).  Anything else -> `cv2.VideoCapture`.  Both feed the
same device path; this is a choice of input format, not a CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import ctypes.util
import glob
import os

import numpy as np

AV_PIX_FMT_YUV420P, AV_PIX_FMT_YUVJ420P = 0, 12
AVMEDIA_TYPE_VIDEO = 0
AVERROR_EOF = -541478725            # FFERRTAG('E','O','F',' ')
AVERROR_EAGAIN = -11


class DecodeUnavailable(RuntimeError):
    pass


# ---- libav through ctypes -----------------------------------------------------------------------------------
class _AVFrameHead(C.Structure):        # the leading, ABI-stable fields of AVFrame (FFmpeg 4 .. 8)
    _fields_ = [("data", C.c_void_p * 8), ("linesize", C.c_int * 8), ("extended_data", C.c_void_p),
                ("width", C.c_int), ("height", C.c_int), ("nb_samples", C.c_int), ("format", C.c_int)]


class _AVPacketHead(C.Structure):       # AVPacket up to stream_index
    _fields_ = [("buf", C.c_void_p), ("pts", C.c_int64), ("dts", C.c_int64), ("data", C.c_void_p), ("size", C.c_int),
                ("stream_index", C.c_int)]


class _AVFormatHead(C.Structure):       # AVFormatContext up to streams
    _fields_ = [("av_class", C.c_void_p), ("iformat", C.c_void_p), ("oformat", C.c_void_p), ("priv_data", C.c_void_p),
                ("pb", C.c_void_p), ("ctx_flags", C.c_int), ("nb_streams", C.c_uint), ("streams", C.POINTER(C.c_void_p))]


class _AVStreamHead(C.Structure):       # AVStream up to codecpar (libavformat >= 60)
    _fields_ = [("av_class", C.c_void_p), ("index", C.c_int), ("id", C.c_int), ("codecpar", C.c_void_p)]


_libav = None


def _find_libs():
    cands = []
    d = os.environ.get("AVD_LIBAV_DIR")
    if d:
        cands.append(d)
    try:
        import cv2
        site = os.path.dirname(os.path.dirname(cv2.__file__))
        cands += [os.path.join(site, n) for n in ("opencv_python_headless.libs", "opencv_python.libs",
                                                  "opencv_contrib_python_headless.libs", "opencv_contrib_python.libs")]
    except Exception:                                        # noqa: BLE001
        pass
    for d in cands:
        got = {}
        for name in ("avutil", "swresample", "avcodec", "avformat"):
            m = sorted(glob.glob(os.path.join(d, f"lib{name}[-.]*so*")))
            if m:
                got[name] = m[0]
        if all(k in got for k in ("avutil", "avcodec", "avformat")):
            return got
    got = {n: ctypes.util.find_library(n) for n in ("avutil", "swresample", "avcodec", "avformat")}
    if all(got.get(k) for k in ("avutil", "avcodec", "avformat")):
        return got
    raise DecodeUnavailable("libavformat / libavcodec / libavutil not found")


def load_libav():
    """Loads the FFmpeg libraries and declares the handful of functions used.  Raises DecodeUnavailable."""
    global _libav
    if _libav is not None:
        if isinstance(_libav, Exception):
            raise _libav
        return _libav
    try:
        paths = _find_libs()
        U = C.CDLL(paths["avutil"], mode=C.RTLD_GLOBAL)
        if paths.get("swresample"):
            C.CDLL(paths["swresample"], mode=C.RTLD_GLOBAL)
        K = C.CDLL(paths["avcodec"], mode=C.RTLD_GLOBAL)
        F = C.CDLL(paths["avformat"], mode=C.RTLD_GLOBAL)
        F.avformat_version.restype = C.c_uint
        if (F.avformat_version() >> 16) < 60:
            raise DecodeUnavailable("libavformat older than 60: AVStream layout not supported by this shim")
        vp, vpp = C.c_void_p, C.POINTER(C.c_void_p)
        F.avformat_open_input.argtypes = [vpp, C.c_char_p, vp, vp]; F.avformat_open_input.restype = C.c_int
        F.avformat_find_stream_info.argtypes = [vp, vp]; F.avformat_find_stream_info.restype = C.c_int
        F.av_find_best_stream.argtypes = [vp, C.c_int, C.c_int, C.c_int, vpp, C.c_int]; F.av_find_best_stream.restype = C.c_int
        F.av_read_frame.argtypes = [vp, vp]; F.av_read_frame.restype = C.c_int
        F.avformat_close_input.argtypes = [vpp]; F.avformat_close_input.restype = None
        K.avcodec_alloc_context3.argtypes = [vp]; K.avcodec_alloc_context3.restype = vp
        K.avcodec_parameters_to_context.argtypes = [vp, vp]; K.avcodec_parameters_to_context.restype = C.c_int
        K.avcodec_open2.argtypes = [vp, vp, vp]; K.avcodec_open2.restype = C.c_int
        K.avcodec_send_packet.argtypes = [vp, vp]; K.avcodec_send_packet.restype = C.c_int
        K.avcodec_receive_frame.argtypes = [vp, vp]; K.avcodec_receive_frame.restype = C.c_int
        K.avcodec_free_context.argtypes = [vpp]; K.avcodec_free_context.restype = None
        K.av_packet_alloc.restype = vp
        K.av_packet_unref.argtypes = [vp]; K.av_packet_unref.restype = None
        K.av_packet_free.argtypes = [vpp]; K.av_packet_free.restype = None
        U.av_frame_alloc.restype = vp
        U.av_frame_unref.argtypes = [vp]; U.av_frame_unref.restype = None
        U.av_frame_free.argtypes = [vpp]; U.av_frame_free.restype = None
        U.av_opt_set_int.argtypes = [vp, C.c_char_p, C.c_int64, C.c_int]; U.av_opt_set_int.restype = C.c_int
        U.av_log_set_level.argtypes = [C.c_int]; U.av_log_set_level.restype = None
        U.av_log_set_level(16)                               # AV_LOG_ERROR
        _libav = (U, K, F)
    except DecodeUnavailable as e:
        _libav = e
        raise
    except Exception as e:                                   # noqa: BLE001 - missing symbol, unloadable library, ...
        _libav = DecodeUnavailable(f"cannot use libav: {e}")
        raise _libav
    return _libav


class LibavSource:
    """Decoder pictures as dense I420 frames: H rows of W luma bytes, then the (H/2)x(W/2) U plane, then V."""

    def __init__(self, path: str, threads: int = 0):
        self.U, self.K, self.F = load_libav()
        U, K, F = self.U, self.K, self.F
        self.fmt_ctx, self.dec, self.pkt, self.frm = C.c_void_p(), C.c_void_p(), None, None
        self.opened, self._draining, self._have, self.fmt = False, False, False, None
        if F.avformat_open_input(C.byref(self.fmt_ctx), os.fsencode(path), None, None) < 0:
            self.fmt_ctx = C.c_void_p()
            return
        if F.avformat_find_stream_info(self.fmt_ctx, None) < 0:
            return
        codec = C.c_void_p()
        self.stream = F.av_find_best_stream(self.fmt_ctx, AVMEDIA_TYPE_VIDEO, -1, -1, C.byref(codec), 0)
        if self.stream < 0 or not codec:
            return
        head = _AVFormatHead.from_address(self.fmt_ctx.value)
        st = _AVStreamHead.from_address(head.streams[self.stream])
        self.dec = C.c_void_p(K.avcodec_alloc_context3(codec))
        if not self.dec or K.avcodec_parameters_to_context(self.dec, st.codecpar) < 0:
            return
        U.av_opt_set_int(self.dec, b"threads", int(threads or min(16, os.cpu_count() or 1)), 0)
        if K.avcodec_open2(self.dec, codec, None) < 0:
            return
        self.pkt, self.frm = C.c_void_p(K.av_packet_alloc()), C.c_void_p(U.av_frame_alloc())
        self.opened = bool(self.pkt and self.frm)
        self.width = self.height = 0

    # -- the reference's loop: grab() every frame, retrieve() the sampled ones ---------------------------------
    def grab(self) -> bool:
        """Decodes the next picture (kept inside libav until retrieve_into)."""
        U, K, F = self.U, self.K, self.F
        if not self.opened:
            return False
        while True:
            rc = K.avcodec_receive_frame(self.dec, self.frm)
            if rc == 0:
                f = _AVFrameHead.from_address(self.frm.value)
                if self.fmt is None:
                    if f.format not in (AV_PIX_FMT_YUV420P, AV_PIX_FMT_YUVJ420P) or (f.width | f.height) & 1:
                        raise DecodeUnavailable(f"pixel format {f.format} / size {f.width}x{f.height}: not 8-bit 4:2:0 with even size")
                    self.fmt = "i420" if f.format == AV_PIX_FMT_YUV420P else "j420"
                    self.width, self.height, self._pixfmt = f.width, f.height, f.format
                elif (f.width, f.height, f.format) != (self.width, self.height, self._pixfmt):
                    raise ValueError("frame size changed mid-stream")
                self._have = True
                return True
            if rc == AVERROR_EOF or (self._draining and rc != AVERROR_EAGAIN):
                return False
            if rc != AVERROR_EAGAIN:
                return False
            # the decoder wants input: next packet of our stream, or start draining at end of file
            while True:
                r = F.av_read_frame(self.fmt_ctx, self.pkt)
                if r < 0:
                    self._draining = True
                    K.avcodec_send_packet(self.dec, None)
                    break
                p = _AVPacketHead.from_address(self.pkt.value)
                if p.stream_index == self.stream:
                    K.avcodec_send_packet(self.dec, self.pkt)
                    K.av_packet_unref(self.pkt)
                    break
                K.av_packet_unref(self.pkt)

    def frame_nbytes(self) -> int:
        return self.width * self.height * 3 // 2

    def retrieve_into(self, dst: np.ndarray) -> bool:
        """Copies the grabbed picture's planes into `dst` (uint8, frame_nbytes() long, e.g. a pinned staging slot)."""
        if not self._have:
            return False
        f = _AVFrameHead.from_address(self.frm.value)
        H, W = self.height, self.width
        off = 0
        for i, (h, w) in enumerate(((H, W), (H // 2, W // 2), (H // 2, W // 2))):
            ls = f.linesize[i]
            if ls < w or not f.data[i]:
                return False
            src = np.ctypeslib.as_array((C.c_uint8 * (ls * h)).from_address(f.data[i])).reshape(h, ls)[:, :w]
            np.copyto(dst[off:off + h * w].reshape(h, w), src)
            off += h * w
        return True

    def release(self):
        if self.frm:
            self.U.av_frame_free(C.byref(self.frm))
        if self.pkt:
            self.K.av_packet_free(C.byref(self.pkt))
        if self.dec:
            self.K.avcodec_free_context(C.byref(self.dec))
        if self.fmt_ctx:
            self.F.avformat_close_input(C.byref(self.fmt_ctx))
        self.opened = False


def open_yuv_source(path: str, cap=None):
    """A LibavSource positioned before the first picture, or None when the 4:2:0 route is not provably equivalent for
    this file (then the caller uses cv2.VideoCapture, like the reference).  `cap`: an already opened cv2.VideoCapture of
    the same file (for its rotation metadata); opened here when not given."""
    if os.environ.get("AVD_DECODE", "auto") == "cv2":
        return None
    try:
        import cv2
        own = cap is None
        if own:
            cap = cv2.VideoCapture(path)
        try:
            if not cap.isOpened():
                return None
            prop = getattr(cv2, "CAP_PROP_ORIENTATION_META", None)
            if prop is not None and (cap.get(prop) or 0.0) != 0.0:
                return None                                   # OpenCV would rotate the BGR frames
        finally:
            if own:
                cap.release()
        src = LibavSource(path)
        if not src.opened:
            src.release()
            return None
        return src
    except DecodeUnavailable:
        return None

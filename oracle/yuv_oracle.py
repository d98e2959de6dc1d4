"""CPU restatement of the colour conversion inside cv2.VideoCapture.retrieve() -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

The reference gets BGR frames from `cap.retrieve()` (video.py:32).  For 8-bit 4:2:0 streams OpenCV's FFmpeg backend
produces them with `sws_scale(..., SWS_BICUBIC)` at unchanged size (cap_ffmpeg_impl.hpp, retrieveFrame), i.e. libswscale's
unscaled yuv420p -> bgr24 special converter.  The algorithm lives in a third-party dependency that is not under
/root/reference: FFmpeg's libswscale (the copy bundled with the opencv-python-headless wheel: libswscale 9.1.100,
FFmpeg 8.0; the reference pins opencv-python-headless>=4.10,<4.11 whose wheel bundles an older FFmpeg with the same
routine).  Restated here from its published source (libswscale/yuv2rgb.c: ff_yuv2rgb_c_init_tables for the coefficients;
libswscale/x86/yuv_2_rgb.asm / yuv2rgb_template.c for the per-pixel SIMD arithmetic, which is what runs on any x86 CPU
with SSSE3) and pinned by tests/golden/golden_yuv.json, produced by calling sws_scale of that very library
(tests/golden/make_golden_yuv.py), and by decoding real MPEG-4 / MJPEG files with cv2.VideoCapture (tests).
"""
from __future__ import annotations

import numpy as np


def _round_to_int16(f: int) -> int:
    r = (f + (1 << 15)) >> 16
    return max(-0x7FFF, min(0x7FFF, r))


def _cdiv(a: int, b: int) -> int:
    """C integer division (truncates toward zero)."""
    q = abs(a) // abs(b)
    return q if (a >= 0) == (b >= 0) else -q


def coefficients(full_range: bool) -> dict:
    """ff_yuv2rgb_c_init_tables with inv_table = ITU601 {104597, 132201, 25675, 53279}, brightness 0, contrast and
    saturation 1.0 (OpenCV never calls sws_setColorspaceDetails)."""
    crv, cbu, cgu, cgv = 104597, 132201, -25675, -53279
    cy, oy = 1 << 16, 0
    if not full_range:
        cy = _cdiv(cy * 255, 219)
        oy = 16 << 16
    else:
        crv, cbu, cgu, cgv = (_cdiv(c * 224, 255) for c in (crv, cbu, cgu, cgv))
    return {"y": _round_to_int16(cy << 13), "vr": _round_to_int16(crv << 13), "ub": _round_to_int16(cbu << 13),
            "vg": _round_to_int16(cgv << 13), "ug": _round_to_int16(cgu << 13), "yo": _round_to_int16(oy * 8)}


def _sat16(a):
    return np.clip(a, -32768, 32767)


def yuv420_to_bgr(Y: np.ndarray, U: np.ndarray, V: np.ndarray, full_range: bool = False) -> np.ndarray:
    """(H,W) luma + (H/2,W/2) chroma planes -> (H,W,3) BGR, the 16-bit lane arithmetic of the SIMD kernel:
    PSLLW 3, PSUB(S)W offset, PMULHW coefficient, PADDSW, PACKUSWB."""
    k = coefficients(full_range)
    H, W = Y.shape
    Yw = Y.astype(np.int32)
    Uw = np.repeat(np.repeat(U.astype(np.int32), 2, 0), 2, 1)[:H, :W]
    Vw = np.repeat(np.repeat(V.astype(np.int32), 2, 0), 2, 1)[:H, :W]
    yw = (Yw << 3) - k["yo"]
    yw = ((yw + 32768) & 0xFFFF) - 32768                     # PSUBW wraps
    yc = (yw * k["y"]) >> 16                                 # PMULHW: high half of the signed product
    uw = _sat16((Uw << 3) - 1024)                            # PSUBSW
    vw = _sat16((Vw << 3) - 1024)
    ub, ug = (uw * k["ub"]) >> 16, (uw * k["ug"]) >> 16
    vr, vg = (vw * k["vr"]) >> 16, (vw * k["vg"]) >> 16
    cg = _sat16(ug + vg)                                     # PADDSW
    B = np.clip(_sat16(yc + ub), 0, 255)                     # PADDSW, PACKUSWB
    G = np.clip(_sat16(yc + cg), 0, 255)
    R = np.clip(_sat16(yc + vr), 0, 255)
    return np.stack([B, G, R], -1).astype(np.uint8)


def i420_planes(frame: np.ndarray, H: int, W: int):
    """Dense I420 frame (1.5*H*W bytes) -> Y, U, V views."""
    f = np.asarray(frame).reshape(-1)
    Y = f[:H * W].reshape(H, W)
    U = f[H * W:H * W + (H // 2) * (W // 2)].reshape(H // 2, W // 2)
    V = f[H * W + (H // 2) * (W // 2):].reshape(H // 2, W // 2)
    return Y, U, V


def i420_to_bgr(frame: np.ndarray, H: int, W: int, full_range: bool = False) -> np.ndarray:
    return yuv420_to_bgr(*i420_planes(frame, H, W), full_range=full_range)

#!/usr/bin/env python
"""Golden vectors for the decoder-side colour conversion, produced by the LIBRARY ITSELF:

    python tests/golden/make_golden_yuv.py        (build container only; writes tests/golden/golden_yuv.json)

cv2.VideoCapture.retrieve() (video.py:32) converts 8-bit 4:2:0 pictures with sws_scale(SWS_BICUBIC) of the libswscale
that ships inside the opencv-python-headless wheel.  This script loads that shared library with ctypes and calls
sws_getContext / sws_scale on seeded random planes (every byte value in every plane, several sizes incl. non-multiple-of-16
widths, both yuv420p and yuvj420p) and on the exhaustive (Y, U, V) cube sampled at all 256 x 256 (U, V) pairs for 256 luma
values.  Stored: sha256 of every BGR output plus a 16 x 16 corner of each for debugging.
Inputs are regenerated from the seeds by the tests.
"""
import ctypes as C
import glob
import hashlib
import json
import os

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
YUV420P, BGR24, YUVJ420P, SWS_BICUBIC = 0, 3, 12, 4


def load_swscale():
    libs = os.path.join(os.path.dirname(os.path.dirname(cv2.__file__)), "opencv_python_headless.libs")
    C.CDLL(glob.glob(libs + "/libavutil-*")[0], mode=C.RTLD_GLOBAL)
    S = C.CDLL(glob.glob(libs + "/libswscale-*")[0])
    S.sws_getContext.restype = C.c_void_p
    S.sws_getContext.argtypes = [C.c_int] * 7 + [C.c_void_p] * 3
    S.sws_scale.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.c_int, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_int)]
    S.sws_scale.restype = C.c_int
    S.swscale_version.restype = C.c_uint
    return S


def sws_bgr(S, Y, U, V, fmt):
    H, W = Y.shape
    ctx = S.sws_getContext(W, H, fmt, W, H, BGR24, SWS_BICUBIC, None, None, None)
    assert ctx
    out = np.zeros((H, W * 3 + 64), np.uint8)              # the SIMD kernel may write past the row end
    src = (C.c_void_p * 4)(Y.ctypes.data, U.ctypes.data, V.ctypes.data, None)
    ss = (C.c_int * 4)(Y.strides[0], U.strides[0], V.strides[0], 0)
    dst = (C.c_void_p * 4)(out.ctypes.data, None, None, None)
    ds = (C.c_int * 4)(out.strides[0], 0, 0, 0)
    assert S.sws_scale(ctx, src, ss, 0, H, dst, ds) == H
    return np.ascontiguousarray(out[:, :W * 3].reshape(H, W, 3))


def planes(seed, H, W):
    rng = np.random.default_rng(seed)
    return (rng.integers(0, 256, (H, W), dtype=np.uint8), rng.integers(0, 256, (H // 2, W // 2), dtype=np.uint8),
            rng.integers(0, 256, (H // 2, W // 2), dtype=np.uint8))


def cube(y0):
    """2 luma values x all (U, V): a 512 x 512 picture whose chroma planes enumerate every (U, V) pair."""
    U, V = np.meshgrid(np.arange(256, dtype=np.uint8), np.arange(256, dtype=np.uint8), indexing="ij")
    Y = np.empty((512, 512), np.uint8)
    Y[0::2, 0::2], Y[0::2, 1::2], Y[1::2, 0::2], Y[1::2, 1::2] = y0, (y0 + 64) & 255, (y0 + 128) & 255, (y0 + 192) & 255
    return Y, np.ascontiguousarray(U), np.ascontiguousarray(V)


def main():
    S = load_swscale()
    v = S.swscale_version()
    out = {"libswscale": f"{v >> 16}.{(v >> 8) & 255}.{v & 255}", "cv2": cv2.__version__, "cases": []}
    for seed, (H, W) in enumerate([(64, 64), (480, 640), (1080, 1920), (480, 854), (50, 66), (720, 1280), (2160, 3840)]):
        for fmt, full in ((YUV420P, False), (YUVJ420P, True)):
            Y, U, V = planes(100 + seed, H, W)
            bgr = sws_bgr(S, Y, U, V, fmt)
            out["cases"].append({"kind": "random", "seed": 100 + seed, "H": H, "W": W, "full_range": full,
                                 "sha256": hashlib.sha256(bgr.tobytes()).hexdigest(), "corner": bgr[:4, :4].tolist()})
    for y0 in range(64):                                  # 64 pictures x 4 luma values each = all 256 luma values
        for fmt, full in ((YUV420P, False), (YUVJ420P, True)):
            Y, U, V = cube(y0)
            bgr = sws_bgr(S, Y, U, V, fmt)
            out["cases"].append({"kind": "cube", "y0": y0, "full_range": full, "sha256": hashlib.sha256(bgr.tobytes()).hexdigest()})
    with open(os.path.join(HERE, "golden_yuv.json"), "w") as f:
        json.dump(out, f, indent=0)
    print("wrote", len(out["cases"]), "cases; libswscale", out["libswscale"])


if __name__ == "__main__":
    main()

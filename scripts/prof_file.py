"""video.analyze(path, meta) on a real file, end to end (decode included): the reference's loop on the CPU (its body restated in
oracle/cv_body.py around the same cv2 calls; /root/reference is not on the GPU box), the drop-in through cv2.VideoCapture (BGR
upload) and through the 4:2:0 route, plus the decode-only floor of both decoders' loops.

    python scripts/prof_file.py [frames] [H] [W] [fourcc]        default: 600 frames of 1080p, mp4v
"""
import os, sys, time, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cv2
import numpy as np
import avd_b200.synth as synth
from avd_b200 import decode, video
from avd_b200.engine import Engine
from oracle import cv_body

n = int(sys.argv[1]) if len(sys.argv) > 1 else 600
H = int(sys.argv[2]) if len(sys.argv) > 2 else 1080
W = int(sys.argv[3]) if len(sys.argv) > 3 else 1920
fourcc = sys.argv[4] if len(sys.argv) > 4 else "mp4v"
path = os.path.join(tempfile.mkdtemp(), "clip." + ("avi" if fourcc == "MJPG" else "mp4"))
clip = synth.Clip(n, H, W, seed=9, fps=30.0)
vw = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*fourcc), 30.0, (W, H))
assert vw.isOpened()
for i in range(n):
    vw.write(clip[i])
vw.release()
meta = {"fps": 30.0, "width": W, "height": H, "duration": n / 30.0}
print(f"{path}: {n} frames {W}x{H} {fourcc}, {os.path.getsize(path) / 1e6:.1f} MB, step 15 -> {len(range(0, n, 15))} samples")


def t(fn, reps=3):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    return (time.perf_counter() - t0) / reps, out


def decode_only_cv2():
    cap = cv2.VideoCapture(path); k = 0
    while cap.grab():
        if k % 15 == 0:
            cap.retrieve()
        k += 1
    cap.release(); return k


def decode_only_libav():
    src = decode.open_yuv_source(path); k = 0
    buf = np.empty(H * W * 3 // 2, np.uint8)
    while src.grab():
        if k % 15 == 0:
            src.retrieve_into(buf)
        k += 1
    src.release(); return k


def reference_cpu():
    cap = cv2.VideoCapture(path); ch = cv_body.SampleChain(); k = 0
    while cap.grab():
        if k % 15 == 0:
            ok, f = cap.retrieve()
            ch.push(f)
        k += 1
    cap.release(); return ch.timeline


eng = Engine(0)
d_cv, _ = t(decode_only_cv2)
d_av, _ = t(decode_only_libav)
os.environ["AVD_DECODE"] = "cv2"
a_cv, r_cv = t(lambda: video.analyze(path, dict(meta), engine=eng))
os.environ["AVD_DECODE"] = "auto"
a_yuv, r_yuv = t(lambda: video.analyze(path, dict(meta), engine=eng))
c_ref, tl = t(reference_cpu, reps=1)
assert r_cv == r_yuv
assert all(abs(a - b) <= 1e-4 * abs(b) + 5e-5 for a, b in zip(r_yuv["timeline"], tl))
print(f"decode only: cv2 grab+retrieve {1e3*d_cv:.0f} ms ({n/d_cv:.0f} source frames/s); libav grab + plane copy {1e3*d_av:.0f} ms ({n/d_av:.0f} frames/s)")
print(f"reference loop on the CPU (cv2 decode + per-sample body, 1 process): {1e3*c_ref:.0f} ms")
print(f"drop-in, cv2.VideoCapture route (BGR upload): {1e3*a_cv:.0f} ms   4:2:0 route: {1e3*a_yuv:.0f} ms   -> {c_ref/a_yuv:.1f}x the reference, "
      f"{100*(a_yuv-d_av)/a_yuv:.0f} % of the call is not decode; results identical on both routes, timeline within tolerance of the CPU loop")

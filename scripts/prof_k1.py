"""Short driver for ncu captures of the per-frame kernel: n frames of HxW noise resident in HBM, a few K1 launches."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from avd_b200.engine import Engine
H, W, n, reps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
eng = Engine(0)
torch.manual_seed(0)
kind = sys.argv[5] if len(sys.argv) > 5 else "noise"
if kind == "noise":
    frames = torch.randint(0, 256, (n, H, W, 3), dtype=torch.uint8, device="cuda")
elif kind == "flat":                     # every INTER_AREA cell mean is an exact integer: no cell near a rounding boundary
    frames = torch.full((n, H, W, 3), 100, dtype=torch.uint8, device="cuda")
else:                                    # "half": every cell mean is exactly k + 0.5: every cell takes the exact-order path
    frames = torch.full((n, H, W, 3), 100, dtype=torch.uint8, device="cuda"); frames[:, :, 1::2, :] = 101
for _ in range(reps):
    rec, th = eng.frames_features(frames)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    rec, th = eng.frames_features(frames)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
eng.profile_read(); eng.profile(True)
for _ in range(reps):
    rec, th = eng.frames_features(frames)
eng.profile(False); kern = eng.profile_read()
k = [v for name, v in kern.items() if name.startswith("k1_")][0]
kms = k["ms"] / k["launches"]
print(f"K1 {H}x{W} n={n} {kind}: {ms:.4f} ms/launch-group, {n*H*W*3/ms/1e6:.1f} GB/s; kernel alone {kms:.4f} ms = {n*H*W*3/kms/1e6:.1f} GB/s "
      f"({n*H*W*3/kms/1e6/65.485:.1f} % of 6548.5)")

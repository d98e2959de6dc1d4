"""Short driver for ncu captures of the per-frame kernel: n frames of HxW noise resident in HBM, a few K1 launches."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from avd_b200.engine import Engine
H, W, n, reps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
eng = Engine(0)
torch.manual_seed(0)
frames = torch.randint(0, 256, (n, H, W, 3), dtype=torch.uint8, device="cuda")
for _ in range(reps):
    rec, th = eng.frames_features(frames)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    rec, th = eng.frames_features(frames)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f"K1 {H}x{W} n={n}: {ms:.4f} ms/launch-group, {n*H*W*3/ms/1e6:.1f} GB/s")

"""End-to-end host path timing: 120 pinned 1080p frames (4:2:0 by default, `bgr` as argv[2]) -> avd_analyze_host(_yuv420)
-> records (bench.py's e2e loop without the Python tail).  Tuning knobs: AVD_HOST_CHUNK, AVD_HOST_K2_MIN, AVD_K2_SPLIT."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from avd_b200.engine import Engine, HostBatch
n = int(sys.argv[1]) if len(sys.argv) > 1 else 120
fmt = sys.argv[2] if len(sys.argv) > 2 else "i420"
H, W = 1080, 1920
eng = Engine(0)
nb = H * W * 3 // 2 if fmt != "bgr" else H * W * 3
pinned = eng.pinned((n, nb))
rng = np.random.default_rng(0)
pinned[:] = rng.integers(0, 256, (1, nb), dtype=np.uint8)
batch = HostBatch(pinned.reshape(n, H, W, 3)) if fmt == "bgr" else HostBatch(pinned, "i420", (H, W))
for _ in range(3):
    eng.analyze_host(batch)
torch.cuda.synchronize()
reps = 20
t0 = time.perf_counter()
for _ in range(reps):
    recs = eng.analyze_host(batch)
dt = (time.perf_counter() - t0) / reps
print(f"fmt={fmt} chunk={os.environ.get('AVD_HOST_CHUNK', '16')} k2min={os.environ.get('AVD_HOST_K2_MIN', '40')} split={os.environ.get('AVD_K2_SPLIT', '-')}: "
      f"{1e3*dt:.3f} ms per {n} frames = {n/dt:.0f} frames/s, {n*nb/dt/1e9:.1f} GB/s over PCIe")

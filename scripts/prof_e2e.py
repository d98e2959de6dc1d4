"""End-to-end host path timing: 120 pinned 1080p frames -> avd_analyze_host -> records (bench.py's e2e loop without the tail)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from avd_b200.engine import Engine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 120
eng = Engine(0)
pinned = eng.pinned((n, 1080, 1920, 3))
rng = np.random.default_rng(0)
pinned[:] = rng.integers(0, 256, (1, 1080, 1920, 3), dtype=np.uint8)
batch = [pinned[i] for i in range(n)]
for _ in range(3):
    eng.analyze_host(batch)
torch.cuda.synchronize()
t0 = time.perf_counter()
reps = 10
for _ in range(reps):
    recs = eng.analyze_host(batch)
dt = (time.perf_counter() - t0) / reps
print(f"chunk={os.environ.get('AVD_HOST_CHUNK', '16')}: analyze_host {1e3*dt:.3f} ms per {n} frames = {n/dt:.0f} frames/s, {n*1080*1920*3/dt/1e9:.1f} GB/s")

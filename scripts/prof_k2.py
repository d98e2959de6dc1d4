"""Short driver for ncu captures of the Farneback kernels: n thumbnails resident in HBM, a few K2 passes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import avd_b200.synth as synth
from avd_b200.engine import Engine
n, reps = int(sys.argv[1]), int(sys.argv[2])
eng = Engine(0)
tex = synth.texture(320, 320 + 4 * n, 5, 0)[..., 1]
thumbs = torch.from_numpy(np.stack([tex[:, 3 * i:3 * i + 320] for i in range(n)])).cuda()
rec = torch.zeros((n, 160), dtype=torch.uint8, device="cuda")
for _ in range(reps):
    eng.flow_stats(thumbs, rec)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    eng.flow_stats(thumbs, rec)
e1.record(); torch.cuda.synchronize()
print(f"K2 n={n}: {e0.elapsed_time(e1)/reps:.4f} ms per pass, {1e3*e0.elapsed_time(e1)/reps/(n-1):.2f} us/pair")
eng.profile_read(); eng.profile(True)
for _ in range(reps):
    eng.flow_stats(thumbs, rec)
eng.profile(False)
for k, v in sorted(eng.profile_read().items(), key=lambda t: -t[1]["ms"]):
    print(f"  {k:18s} {1e3*v['ms']/reps:8.1f} us/pass  {v['launches']//reps} launches")

"""Per-kernel breakdown of the generic per-frame path (any shape): python scripts/prof_generic.py H W n"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from avd_b200.engine import Engine
H, W, n = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
eng = Engine(0)
frames = torch.randint(0, 256, (n, H, W, 3), dtype=torch.uint8, device="cuda")
for _ in range(3):
    eng.frames_features(frames)
torch.cuda.synchronize()
eng.profile_read(); eng.profile(True)
for _ in range(5):
    eng.frames_features(frames)
eng.profile(False)
k = eng.profile_read()
tot = sum(v["ms"] for v in k.values()) / 5
print(f"{H}x{W} n={n}: {tot:.3f} ms per call = {n*H*W*3/tot/1e6:.0f} GB/s")
for name, v in sorted(k.items(), key=lambda kv: -kv[1]["ms"]):
    print(f"  {name:14s} {1e3*v['ms']/5:8.1f} us  {v['launches']//5} launches")

"""Small K1/K2 run through every per-frame kernel path (streaming, banded, generic; boundary-cell fix-up) and one short
host pipeline -- a quick crash check after kernel changes (compute-sanitizer is closed on this pool)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from avd_b200.engine import Engine
eng = Engine(0)
for (H, W) in [(720, 1280), (1080, 1920), (1440, 2560), (2160, 3840), (648, 1344), (480, 640), (333, 777)]:
    fr = torch.randint(0, 256, (2, H, W, 3), dtype=torch.uint8, device="cuda")
    half = torch.full((1, H, W, 3), 100, dtype=torch.uint8, device="cuda"); half[:, :, 1::2, :] = 101
    for mode in (0, 2, 3, 1):
        eng.set_k1_mode(mode)
        eng.frames_features(fr, want_area32=True)
    eng.set_k1_mode(0)
    eng.frames_features(half)
    torch.cuda.synchronize()
    print(H, W, "ok", flush=True)
frames = np.random.default_rng(0).integers(0, 256, (5, 720, 1280, 3), dtype=np.uint8)
print(eng.analyze_host(list(frames))["ham_prev"])
eng.close()

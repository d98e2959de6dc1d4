"""The any-width per-frame kernel against the generic path on extreme aspect ratios (1-row bands, 128-row bands, 2-group rows)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np, torch
from avd_b200.engine import Engine
from test_host_logic import _plan
eng=Engine(0); bad=0
rng=np.random.default_rng(3)
for (H,W) in [(4096,64),(4096,33),(64,4096),(35,4096),(2048,100),(4095,47),(1000,4090),(720,1281),(719,1279),(32,4096),(4096,32),(500,2049)]:
    fr=torch.from_numpy(rng.integers(0,256,(3,H,W,3),dtype=np.uint8)).cuda()
    outs=[]
    for mode in (3,1):
        eng.set_k1_mode(mode)
        rec,th,ar=eng.frames_features(fr,want_area32=True)
        outs.append((rec.cpu().numpy().tobytes(),th.cpu().numpy().tobytes(),ar.cpu().numpy().tobytes()))
    eng.set_k1_mode(0)
    ok=outs[0]==outs[1]; bad+=not ok
    print(f"{W}x{H} path={_plan(H,W).k1_path}:", "ok" if ok else "MISMATCH", flush=True)
print("guards", eng.check_guards(), "bad", bad)

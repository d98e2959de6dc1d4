"""Two-frames-per-CTA form of the streaming kernel (W = 1280): byte for byte against the generic path for odd / even frame
counts and several heights, then its kernel-only time against the one-frame form (AVD_K1_NOPAIR=1 in a second process)."""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import avd_b200.synth as synth
from avd_b200.engine import Engine

eng = Engine(0)
if "--time" in sys.argv:
    H, W, n = int(sys.argv[2]), 1280, int(sys.argv[3])
    frames = torch.randint(0, 256, (n, H, W, 3), dtype=torch.uint8, device="cuda")
    for _ in range(3):
        eng.frames_features(frames)
    torch.cuda.synchronize(); eng.profile_read(); eng.profile(True)
    for _ in range(5):
        eng.frames_features(frames)
    eng.profile(False); p = eng.profile_read()["k1_fused"]
    ms = p["ms"] / p["launches"]
    print(f"1280x{H} n={n} NOPAIR={os.environ.get('AVD_K1_NOPAIR', '0')}: {ms:.4f} ms = {n * H * W * 3 / ms / 1e6:.0f} GB/s = {n * H * W * 3 / ms / 1e6 / 65.485:.1f} %")
    sys.exit(0)
bad = 0
for H in (720, 960, 641, 1000, 4096):
    clip = synth.Clip(40, H, 1280, seed=7, events=True)
    for n in (1, 2, 5, 8):
        fr = torch.from_numpy(np.stack([clip[i] for i in (0, 3, 12, 13, 20, 25, 31, 39)[:n]])).cuda()
        outs = []
        for mode in (0, 1):
            eng.set_k1_mode(mode)
            rec, th, ar = eng.frames_features(fr, want_area32=True)
            outs.append((rec.cpu().numpy().tobytes(), th.cpu().numpy().tobytes(), ar.cpu().numpy().tobytes()))
        eng.set_k1_mode(0)
        ok = outs[0] == outs[1]
        if not ok:
            which = [k for k, (a, b) in zip(("rec", "thumbs", "area"), zip(outs[0], outs[1])) if a != b]
            print(f"1280x{H} n={n}: MISMATCH {which}", flush=True)
        bad += not ok
print("parity:", "ok" if not bad else f"{bad} mismatches", "guards:", eng.check_guards(), flush=True)
for H, n in ((720, 256), (720, 120), (960, 256)):
    for nop in ("0", "1"):
        subprocess.run([sys.executable, __file__, "--time", str(H), str(n)], env=dict(os.environ, AVD_K1_NOPAIR=nop))
sys.exit(1 if bad else 0)

"""Any-width per-frame kernel (k1_rows.cu, K1 mode 3) against the generic path (mode 1), output by output, then its
kernel-only time at 256 frames per launch next to the banded kernel where that one applies.

    python scripts/check_k1_rows.py [--time-only]
"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import avd_b200.synth as synth
from avd_b200.engine import Engine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6548.5
eng = Engine(0)


def run(fr, mode):
    eng.set_k1_mode(mode)
    rec, th, ar = eng.frames_features(fr, want_area32=True)
    eng.set_k1_mode(0)
    return eng.records_to_numpy(rec), th.cpu().numpy(), ar.cpu().numpy()


def check():
    bad = 0
    shapes = [(480, 854), (480, 640), (768, 1366), (333, 777), (144, 256), (640, 640), (32, 32), (33, 35), (641, 1283), (1080, 1918),
              (100, 150), (540, 960), (360, 640), (96, 128), (270, 480), (720, 1280), (1080, 1920), (648, 704), (120, 203), (47, 33)]
    for (H, W) in shapes:
        clip = synth.Clip(12, H, W, seed=15, events=False, pan=3)
        fr = torch.from_numpy(np.stack([clip[i] for i in (0, 1, 5, 5, 11)])).cuda()
        r1, t1, a1 = run(fr, 1)
        r3, t3, a3 = run(fr, 3)
        msgs = []
        for k in ("lap_sum", "lap_sumsq"):
            if not np.array_equal(r1[k], r3[k]): msgs.append(f"{k}: {r3[k].tolist()} != {r1[k].tolist()}")
        if not np.array_equal(r1["hash"], r3["hash"]): msgs.append("hash")
        if not np.array_equal(a1, a3):
            w = np.argwhere(a1 != a3); msgs.append(f"area32: {len(w)} cells, first {w[0].tolist()} {a3[tuple(w[0])]} != {a1[tuple(w[0])]}")
        if not np.array_equal(t1, t3):
            w = np.argwhere(t1 != t3); msgs.append(f"thumb: {len(w)} px, first {w[0].tolist()} {t3[tuple(w[0])]} != {t1[tuple(w[0])]}, rows {sorted(set(w[:, 1].tolist()))[:8]} cols {sorted(set(w[:, 2].tolist()))[:8]}")
        if r1.tobytes() != r3.tobytes() and not msgs: msgs.append("records differ elsewhere")
        print(f"{W}x{H}: " + ("ok" if not msgs else "MISMATCH " + "; ".join(msgs)), flush=True)
        bad += bool(msgs)
    # unaligned views: base pointer and pitch at every byte phase
    for off in (1, 2, 3, 5, 8, 13):
        H, W = 120, 203
        f = synth.noise_frames(2, H, W, 19)
        big = torch.zeros((2, H, W * 3 + 16), dtype=torch.uint8, device="cuda")
        big[:, :, off:off + W * 3] = torch.from_numpy(f.reshape(2, H, W * 3)).cuda()
        view = big[:, :, off:off + W * 3].unflatten(2, (W, 3))
        r1, t1, a1 = run(view, 1)
        r3, t3, a3 = run(view, 3)
        ok = r1.tobytes() == r3.tobytes() and np.array_equal(t1, t3) and np.array_equal(a1, a3)
        print(f"unaligned view offset {off}: " + ("ok" if ok else "MISMATCH"), flush=True)
        bad += not ok
    print("guards:", eng.check_guards())
    return bad


def timing():
    print("| shape | mode | kernel | ms / 256 frames | GB/s | of measured peak |")
    print("|---|---|---|---|---|---|")
    shapes = [(480, 854), (480, 640), (768, 1366), (540, 960), (360, 640), (720, 1280), (1080, 1918)]
    modes = (0, 3)
    if "--crossover" in sys.argv:      # banded (2) against any-width (3) on shapes both take
        shapes = [(576, 1024), (544, 1280), (600, 1920), (1080, 1936), (720, 1296), (800, 1600), (480, 1280)]
        modes = (2, 3)
    for (H, W) in shapes:
        n = 256
        frames = torch.randint(0, 256, (n, H, W, 3), dtype=torch.uint8, device="cuda")
        for mode in modes:
            eng.set_k1_mode(mode)
            for _ in range(3):
                eng.frames_features(frames)
            torch.cuda.synchronize()
            eng.profile_read(); eng.profile(True)
            for _ in range(5):
                eng.frames_features(frames)
            eng.profile(False)
            prof = eng.profile_read()
            eng.set_k1_mode(0)
            name = next((k for k in ("k1_fused", "k1_banded", "k1_rows") if k in prof), None)
            if name is None:
                print(f"| {W}x{H} | {mode} | generic | - | - | - |"); continue
            ms = prof[name]["ms"] / prof[name]["launches"]
            gbs = n * H * W * 3 / ms / 1e6
            print(f"| {W}x{H} | {mode} | {name} | {ms:.4f} | {gbs:.0f} | {100 * gbs / PEAK:.1f} % |", flush=True)
        del frames


if __name__ == "__main__":
    bad = 0 if "--time-only" in sys.argv else check()
    timing()
    sys.exit(1 if bad else 0)

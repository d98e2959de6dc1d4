"""BASELINE configs[4]: resolution sweep of the per-frame kernel (K1) at 1 GPU vs the HBM roofline and the host CPU path.

For each resolution: n frames of noise resident in HBM (worst case: incompressible, maximum Laplacian energy), K1 timed with
CUDA events over `reps` launches after warm-up; the CPU side times the same per-frame library calls the reference makes for
these features (cvtColor x3, INTER_AREA 32x32 + mean test, INTER_LINEAR 320x320, Laplacian(CV_64F).var(); video.py:5-8,43,51-52)
on one core and on all cores.  Prints a markdown table and one JSON line."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from avd_b200.engine import Engine

PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"] \
    if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) else 6650.0


def cpu_frame(f):
    import cv2
    g = cv2.resize(cv2.cvtColor(f, cv2.COLOR_BGR2GRAY), (32, 32), interpolation=cv2.INTER_AREA)
    h = (g >= g.mean())
    t = cv2.resize(cv2.cvtColor(f, cv2.COLOR_BGR2GRAY), (320, 320))
    v = cv2.Laplacian(cv2.cvtColor(f, cv2.COLOR_BGR2GRAY), cv2.CV_64F).var()
    return h.sum() + int(t[0, 0]) + v


def _cpu_worker(args):
    import cv2
    cv2.setNumThreads(1)
    H, W, n = args
    f = np.random.default_rng(0).integers(0, 256, (H, W, 3), dtype=np.uint8)
    cpu_frame(f)
    t0 = time.perf_counter()
    for _ in range(n):
        cpu_frame(f)
    return n / (time.perf_counter() - t0)


def main():
    import multiprocessing as mp
    cores = len(os.sched_getaffinity(0))
    shapes = [(360, 640), (480, 854), (480, 640), (540, 960), (768, 1366), (720, 1280), (1080, 1920), (1440, 2560), (2160, 3840)]
    cpu = {}
    with mp.get_context("fork").Pool(cores) as pool:       # before CUDA is touched
        for (H, W) in shapes:
            n = max(2, int(3e8 / (H * W * 3)) // 8)
            cpu[(H, W)] = (_cpu_worker((H, W, n)), sum(pool.map(_cpu_worker, [(H, W, n)] * cores)))
    eng = Engine(0)
    rows = []
    for (H, W) in shapes:
        n = 256
        frames = torch.randint(0, 256, (n, H, W, 3), dtype=torch.uint8, device="cuda")
        for _ in range(3):
            eng.frames_features(frames)
        torch.cuda.synchronize()
        eng.profile_read(); eng.profile(True)
        reps = 5
        for _ in range(reps):
            eng.frames_features(frames)
        eng.profile(False)
        prof = eng.profile_read()
        k1 = prof.get("k1_fused") or prof.get("k1_banded") or prof.get("k1_rows") or {"ms": sum(prof[k]["ms"] for k in ("gray", "lap", "area_h", "area_v", "area_int", "thumb") if k in prof)}
        ms = k1["ms"] / reps
        gbs = n * H * W * 3 / ms / 1e6
        rows.append({"H": H, "W": W, "frames_per_launch": n, "path": "streaming" if "k1_fused" in prof else ("banded" if "k1_banded" in prof else ("any-width" if "k1_rows" in prof else "generic")),
                     "ms_per_launch": ms, "frames_per_s": n / ms * 1e3, "GBps": gbs, "frac_measured_peak": gbs / PEAK, "frac_8TBps": gbs / 8000.0,
                     "cpu_1core_fps": cpu[(H, W)][0], "cpu_allcores_fps": cpu[(H, W)][1], "cpu_cores": cores})
        del frames
        torch.cuda.empty_cache()
    print("| resolution | path | ms / 256 frames | frames/s | GB/s | of measured %.0f GB/s | of 8 TB/s | CPU 1 core fps | CPU %d cores fps |" % (PEAK, cores))
    print("|---|---|---|---|---|---|---|---|---|")
    for r in rows:
        print(f"| {r['W']}x{r['H']} | {r['path']} | {r['ms_per_launch']:.3f} | {r['frames_per_s']:.0f} | {r['GBps']:.0f} | {100*r['frac_measured_peak']:.1f} % | "
              f"{100*r['frac_8TBps']:.1f} % | {r['cpu_1core_fps']:.1f} | {r['cpu_allcores_fps']:.0f} |")
    print(json.dumps({"sweep_k1": rows}))


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Goldens for BASELINE configs[2] and configs[3] at FULL size, produced by the reference's own code:

    python tests/golden/make_golden_configs23.py     (build container only; writes golden_4k_full.json, golden_clips256.json)

configs[2]: the 4K 60 fps 30 s seed-2 clip (1800 frames -> 60 samples) through the unmodified video.analyze() (video.py:10-83)
            with every library intermediate recorded (same proxy as make_golden.py).
configs[3]: 256 short 720p clips (seeds 1000..1255, 150 frames at 30 fps -> 10 samples, duration 5 s) through the unmodified
            video.analyze() and fusion.fuse() (neutral audio, hints of a 2.5 Mbit/s 720p stream): the returned dicts.
bench.py's sub-results for these configs and the GPU tests compare against them.
"""
import copy
import json
import os
import sys
from concurrent.futures import ProcessPoolExecutor

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as MG  # noqa: E402


def config2():
    H, W, FPS, N, SEED = 2160, 3840, 60.0, 1800, 2
    clip = MG.synth.Clip(N, H, W, SEED, fps=FPS)
    meta = {"fps": FPS, "width": W, "height": H, "duration": N / FPS}
    out, log = MG.run_reference(clip, FPS, meta)
    hashes = [MG.pack_hash(a) for a in log["area32"]]
    fst = []
    for fl in log["flow"]:
        mag = np.sqrt(fl[..., 0] ** 2 + fl[..., 1] ** 2)
        fst.append([float(np.mean(mag)), float(np.var(mag))])
    return {"versions": {"cv2": MG.cv2.__version__, "numpy": np.__version__}, "n_frames": N, "H": H, "W": W, "seed": SEED, "fps": FPS,
            "step": 30, "n_samples": len(hashes), "meta": meta, "result": out,
            "hash_words": [h.tolist() for h in hashes],
            "ham": [int(np.unpackbits((hashes[i] ^ hashes[i - 1]).view(np.uint8)).sum()) for i in range(1, len(hashes))],
            "thumb_sha": [MG.sha(t) for t in log["thumb"]],
            "lap_var": log["lap_var"], "lap_sum": log["lap_sum"], "lap_sumsq": log["lap_sumsq"], "flow_stats": fst}


def one_clip(seed):
    H, W, FPS, N = 720, 1280, 30.0, 150
    clip = MG.synth.Clip(N, H, W, seed, fps=FPS)
    meta = {"fps": FPS, "width": W, "height": H, "duration": 5.0}
    out, log = MG.run_reference(clip, FPS, meta)
    hints = MG.ref_hints.compute_hints(dict(meta, bit_rate=2_500_000), "")
    fused = MG.ref_fusion.fuse({"scores": {}, "flags_audio": {}, "timeline": [0.5] * len(out["timeline"])}, copy.deepcopy(out), hints)
    return {"seed": seed, "video": out, "fused": fused, "lap_sum0": log["lap_sum"][0], "hash0": MG.pack_hash(log["area32"][0]).tolist()}


def main():
    g2 = config2()
    with open(os.path.join(HERE, "golden_4k_full.json"), "w") as f:
        json.dump(g2, f, indent=0)
    print("configs[2]:", g2["n_samples"], "samples")
    with ProcessPoolExecutor(8) as ex:
        clips = list(ex.map(one_clip, range(1000, 1256)))
    hints = MG.ref_hints.compute_hints({"fps": 30.0, "width": 1280, "height": 720, "duration": 5.0, "bit_rate": 2_500_000}, "")
    with open(os.path.join(HERE, "golden_clips256.json"), "w") as f:
        json.dump({"versions": {"cv2": MG.cv2.__version__, "numpy": np.__version__}, "H": 720, "W": 1280, "fps": 30.0, "n_frames": 150,
                   "step": 15, "hints_meta_bit_rate": 2_500_000, "hints": hints, "clips": clips}, f, indent=0)
    print("configs[3]:", len(clips), "clips; labels", {c["fused"]["result"]["label"] for c in clips})


if __name__ == "__main__":
    main()

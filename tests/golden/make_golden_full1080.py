#!/usr/bin/env python
"""Golden for BASELINE configs[1] at FULL size: the reference's unmodified video.analyze() (video.py:10-83) and
fusion.fuse() (fusion.py:16-109) on the whole 1080p 30 fps 60 s seed-1 synthetic clip (1800 frames -> 120 samples).

    python tests/golden/make_golden_full1080.py      (build container only; writes tests/golden/golden_1080p_full.json)

Same mechanism as make_golden.py (whose helpers are reused): the reference module's `cv2` is replaced by a proxy that
serves the in-memory frames through a VideoCapture look-alike and records every library result the analyzer obtains.
This is the clip bench.py analyses, so the bench line can carry a parity verdict against the reference itself.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as MG  # noqa: E402  (loads the reference modules by file path)

H, W, FPS, N, SEED = 1080, 1920, 30.0, 1800, 1


def main():
    clip = MG.synth.Clip(N, H, W, SEED, fps=FPS)
    meta = {"fps": FPS, "width": W, "height": H, "duration": N / FPS}
    out, log = MG.run_reference(clip, FPS, meta)
    hashes = [MG.pack_hash(a) for a in log["area32"]]
    fst = []
    for fl in log["flow"]:
        mag = np.sqrt(fl[..., 0] ** 2 + fl[..., 1] ** 2)
        fst.append([float(np.mean(mag)), float(np.var(mag))])
    hmeta = dict(meta, bit_rate=8_000_000)
    hints = MG.ref_hints.compute_hints(dict(hmeta), "")
    import copy
    tl = len(out["timeline"])
    fused = MG.ref_fusion.fuse({"scores": {}, "flags_audio": {}, "timeline": [0.5] * tl}, copy.deepcopy(out), hints)
    G = {"versions": {"cv2": MG.cv2.__version__, "numpy": np.__version__},
         "n_frames": N, "H": H, "W": W, "seed": SEED, "fps": FPS, "step": 15, "n_samples": len(hashes), "meta": meta,
         "hints_meta": hmeta, "hints": hints,
         "input_sha_first8": MG.sha(np.stack([clip[i] for i in range(0, 8 * 15, 15)])),
         "result": out, "fused_with_neutral_audio": fused,
         "hash_words": [h.tolist() for h in hashes],
         "ham": [int(np.unpackbits((hashes[i] ^ hashes[i - 1]).view(np.uint8)).sum()) for i in range(1, len(hashes))],
         "thumb_sha": [MG.sha(t) for t in log["thumb"]],
         "lap_var": log["lap_var"], "lap_sum": log["lap_sum"], "lap_sumsq": log["lap_sumsq"], "flow_stats": fst}
    p = os.path.join(HERE, "golden_1080p_full.json")
    with open(p, "w") as f:
        json.dump(G, f, indent=0)
    print("wrote", p, os.path.getsize(p), "bytes;", len(hashes), "samples; ai_score", fused["result"])


if __name__ == "__main__":
    main()

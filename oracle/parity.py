"""Parity verdicts for bench.py and the tests -- TEST/BENCH INFRASTRUCTURE, NOT PRODUCT CODE.

Compares the records the GPU path produced with (a) the outputs of the reference's per-sample body on the CPU
(oracle/cv_body.py: the same cv2/NumPy calls as video.py:36-57) for the same frames and (b) the committed golden of the
reference's own analyze() (tests/golden/golden_1080p_full.json).  Gates are BASELINE.json's: integer work bit-exact,
floats |d| <= 1e-4 |ref| + floor (floor 1e-6; 5e-5 px for the flow statistics of static pairs, tests/conftest.py),
pre-rounding fused score |d| <= 1e-3.
"""
from __future__ import annotations

import numpy as np

RTOL, ATOL, ATOL_STATIC, STATIC_MEAN = 1e-4, 1e-6, 5e-5, 1e-2
REC_HAS_PREV, REC_HAS_FLOW = 1, 2


def _textures(recs, npix):
    s, ss, n = recs["lap_sum"].astype(np.float64), recs["lap_sumsq"].astype(np.float64), float(npix)
    return (ss - s * s / n) / n


def _flow_check(got_m, got_v, ref_m, ref_v):
    """(ok, max relative error over moving pairs, max absolute error over static pairs)."""
    ok, max_rel, max_abs_static = True, 0.0, 0.0
    for gm, gv, rm, rv in zip(got_m, got_v, ref_m, ref_v):
        floor = ATOL_STATIC if abs(rm) < STATIC_MEAN else ATOL
        for g, r in ((gm, rm), (gv, rv)):
            d = abs(float(g) - float(r))
            ok &= d <= RTOL * abs(r) + floor
            if abs(rm) < STATIC_MEAN:
                max_abs_static = max(max_abs_static, d)
            elif abs(r) > 1e-3:
                max_rel = max(max_rel, d / abs(r))
    return bool(ok), max_rel, max_abs_static


def records_vs_cpu_body(recs: np.ndarray, cpu: dict, npix: int, first_has_prev: bool = False) -> dict:
    """recs: GPU records of consecutive samples; cpu: cv_body outputs for the same samples (textures, flow_means,
    flow_vars, dup, timeline [, hash_words, hams, lap_sums]).  first_has_prev: the chunk was seeded by a halo sample."""
    n = len(recs)
    tex_g, tex_c = _textures(recs, npix), np.asarray(cpu["textures"], np.float64)
    tex_rel = float(np.max(np.abs(tex_g - tex_c) / np.maximum(np.abs(tex_c), 1e-300))) if n else 0.0
    has_flow = (recs["flags"] & REC_HAS_FLOW) != 0
    fm_g, fv_g = recs["flow_mean"][has_flow], recs["flow_var"][has_flow]
    flows_match = len(fm_g) == len(cpu["flow_means"]) == (n if first_has_prev else max(n - 1, 0))
    f_ok, f_rel, f_abs = _flow_check(fm_g, fv_g, cpu["flow_means"], cpu["flow_vars"]) if flows_match else (False, float("inf"), float("inf"))
    dup_g = int(np.count_nonzero(((recs["flags"] & REC_HAS_PREV) != 0) & (recs["ham_prev"] == 0)))
    int_exact = dup_g == int(cpu["dup"]) and tex_rel <= 1e-12      # texture is a function of the exact integer sums
    if "hash_words" in cpu:
        int_exact &= [r.tolist() for r in recs["hash"]] == [np.asarray(h).tolist() for h in cpu["hash_words"]]
        hams = [int(h) for h in recs["ham_prev"]]
        int_exact &= hams[(0 if first_has_prev else 1):] == [int(h) for h in cpu["hams"]][(0 if first_has_prev else 1):]
        int_exact &= [(int(a), int(b)) for a, b in zip(recs["lap_sum"], recs["lap_sumsq"])] == [tuple(x) for x in cpu["lap_sums"]]
    return {"samples": n, "int_exact": bool(int_exact), "texture_max_rel": tex_rel, "flow_max_rel": f_rel,
            "flow_static_max_abs": f_abs, "ok": bool(int_exact and f_ok)}


def timeline_vs_cpu(timeline, cpu_timeline) -> dict:
    n = min(len(timeline), len(cpu_timeline))
    d = [abs(a - b) for a, b in zip(timeline[:n], cpu_timeline[:n])]
    ok = all(x <= RTOL * abs(b) + ATOL_STATIC for x, b in zip(d, cpu_timeline[:n]))
    return {"timeline_max_abs": max(d) if d else 0.0, "ok": bool(ok)}


def records_vs_reference_golden(recs: np.ndarray, g: dict) -> dict:
    """All samples of BASELINE configs[1] against the reference's own analyze() run (golden_1080p_full.json)."""
    n = len(recs)
    int_exact = (n == g["n_samples"] and [r.tolist() for r in recs["hash"]] == g["hash_words"]
                 and [int(x) for x in recs["ham_prev"][1:]] == g["ham"]
                 and [int(x) for x in recs["lap_sum"]] == g["lap_sum"] and [int(x) for x in recs["lap_sumsq"]] == g["lap_sumsq"])
    ref_m, ref_v = [m for m, _ in g["flow_stats"]], [v for _, v in g["flow_stats"]]
    f_ok, f_rel, f_abs = _flow_check(recs["flow_mean"][1:], recs["flow_var"][1:], ref_m, ref_v) if n == g["n_samples"] else (False, float("inf"), float("inf"))
    tex = _textures(recs, g["H"] * g["W"])
    tex_rel = float(np.max(np.abs(tex - np.asarray(g["lap_var"])) / np.maximum(np.abs(g["lap_var"]), 1e-300))) if n == g["n_samples"] else float("inf")
    return {"samples": n, "int_exact": bool(int_exact), "texture_max_rel": tex_rel, "flow_max_rel": f_rel,
            "flow_static_max_abs": f_abs, "ok": bool(int_exact and f_ok and tex_rel <= 1e-12)}


def result_vs_reference_golden(result: dict, fused_raw: float, ref_result: dict, ref_fused_raw: float, ref_fused: dict, fused: dict) -> dict:
    tl = timeline_vs_cpu(result["timeline"], ref_result["timeline"])
    same_len = len(result["timeline"]) == len(ref_result["timeline"])
    score_delta = abs(float(fused_raw) - float(ref_fused_raw))
    ok = (tl["ok"] and same_len and score_delta <= 1e-3 and fused["result"]["label"] == ref_fused["result"]["label"]
          and fused["peaks"] == ref_fused["peaks"] and len(fused["timeline_binned"]) == len(ref_fused["timeline_binned"]))
    return {"timeline_max_abs": tl["timeline_max_abs"], "score_delta": score_delta, "label": fused["result"]["label"],
            "label_equal": fused["result"]["label"] == ref_fused["result"]["label"], "ok": bool(ok)}


def chain_consistent(recs: np.ndarray) -> bool:
    """A checksum of checksums for gathered shards: every record's Hamming distance must be the distance between its own
    hash and its predecessor's -- across shard boundaries this holds only if the halo sample seeded the chain correctly."""
    h = recs["hash"]
    for i in range(1, len(recs)):
        d = int(np.unpackbits((h[i] ^ h[i - 1]).view(np.uint8)).sum())
        if int(recs["ham_prev"][i]) != d or not (int(recs["flags"][i]) & REC_HAS_PREV) or not (int(recs["flags"][i]) & REC_HAS_FLOW):
            return False
    return len(recs) == 0 or int(recs["ham_prev"][0]) == -1


def records_vs_golden_range(recs: np.ndarray, g: dict, lo: int, first_has_prev: bool) -> dict:
    """Records of samples [lo, lo + len(recs)) of a clip against the reference's golden for the whole clip (any clip golden made
    by tests/golden/make_golden_*.py): integer fields exact, flow statistics in tolerance."""
    n = len(recs)
    hi = lo + n
    int_exact = ([r.tolist() for r in recs["hash"]] == g["hash_words"][lo:hi]
                 and [int(x) for x in recs["lap_sum"]] == g["lap_sum"][lo:hi] and [int(x) for x in recs["lap_sumsq"]] == g["lap_sumsq"][lo:hi])
    k0 = 0 if (first_has_prev and lo > 0) else 1                      # first record of the range that has a predecessor
    int_exact &= [int(x) for x in recs["ham_prev"][k0:]] == g["ham"][lo + k0 - 1:hi - 1]
    ref = g["flow_stats"][lo + k0 - 1:hi - 1]
    f_ok, f_rel, f_abs = _flow_check(recs["flow_mean"][k0:], recs["flow_var"][k0:], [m for m, _ in ref], [v for _, v in ref])
    return {"samples": n, "int_exact": bool(int_exact), "flow_max_rel": f_rel, "flow_static_max_abs": f_abs, "ok": bool(int_exact and f_ok)}


def video_dict_vs_golden(v: dict, ref: dict) -> bool:
    """What video.analyze returns against the reference's dict for the same clip."""
    if len(v["timeline"]) != len(ref["timeline"]) or not timeline_vs_cpu(v["timeline"], ref["timeline"])["ok"]:
        return False
    a, b = v["summary"], ref["summary"]
    if any(a[k] != b[k] for k in ("dup_density", "scene_change_rate", "w", "h", "fps")):
        return False
    return all(abs(a[k] - b[k]) <= RTOL * abs(b[k]) + ATOL_STATIC for k in ("flow_mean", "flow_var")) and abs(a["texture_var"] - b["texture_var"]) <= 1e-9 * abs(b["texture_var"]) + 1e-12


def fused_vs_golden(f: dict, ref: dict) -> bool:
    """fusion.fuse output against the reference's: label, peaks, binned timeline; ai_score / confidence may differ by one unit
    in the last rounded digit only when the pre-rounding value sits on a x.xx5 boundary."""
    return (f["result"]["label"] == ref["result"]["label"] and f["peaks"] == ref["peaks"]
            and abs(f["result"]["ai_score"] - ref["result"]["ai_score"]) <= 0.01 + 1e-12
            and abs(f["result"]["confidence"] - ref["result"]["confidence"]) <= 0.01 + 1e-12
            and len(f["timeline_binned"]) == len(ref["timeline_binned"])
            and all(abs(x - y) <= 1e-3 for x, y in zip(f["timeline_binned"], ref["timeline_binned"])))

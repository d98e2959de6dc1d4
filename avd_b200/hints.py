"""Metadata hints (heuristics_v2.py:1-18 of the reference): the scalar inputs fusion needs besides the
video summary.  Pure host arithmetic."""
from __future__ import annotations

_CLASSES = ((0.04, "very_heavy"), (0.08, "heavy"), (0.15, "normal"))


def compression_class(bpp: float) -> str:
    for bound, name in _CLASSES:
        if bpp <= bound:
            return name
    return "light"


def compute_hints(meta: dict, path: str = "") -> dict:
    w = meta.get("width") or 0
    h = meta.get("height") or 0
    fps = meta.get("fps") or 0.0
    br = meta.get("bit_rate") or 0
    px_per_s = (w * h * fps) if w and h and fps else 0.0
    bpp = float(br) / max(1.0, px_per_s)
    return {"w": w, "h": h, "fps": fps, "br": br, "bpp": round(bpp, 5), "compression": compression_class(bpp),
            "video_has_signal": (w * h) > 0 and fps > 0, "dup_avg": 0.0}

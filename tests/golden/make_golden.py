"""Generate tests/golden/golden.json + golden_arrays.npz by EXECUTING THE REFERENCE'S OWN CODE.

Run in the build container only (needs /root/reference, cv2, numpy):

    python tests/golden/make_golden.py

How the reference is executed, unmodified:
* /root/reference/app/analyzers/{video,fusion,heuristics_v2}.py are loaded by file path (the package
  __init__ imports audio.py -> soundfile, which is not installed).
* video.analyze(path, meta) opens its input with cv2.VideoCapture.  To feed it *exactly the frames the
  GPU path will see*, the module's ``cv2`` global is replaced by a proxy that forwards every attribute
  to the real cv2 except VideoCapture, which serves in-memory frames; the proxy also records the
  intermediate results of the library calls the analyzer makes (resize -> 32x32 / 320x320, Laplacian,
  calcOpticalFlowFarneback) so each stage gets its own golden.  One case additionally goes through a
  real lossless FFV1 .avi and the real cv2.VideoCapture to pin the decode seam.

Inputs are NOT stored: they are regenerated from seeds by avd_b200/synth.py (integer-only,
machine independent); a sha256 of every generated input is stored so generator drift is detected.
Library versions used are recorded in the fixture.
"""
import hashlib
import importlib.util
import json
import os
import sys
import tempfile

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference/app/analyzers"
OUT = os.path.dirname(os.path.abspath(__file__))


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


synth = _load(os.path.join(ROOT, "avd_b200", "synth.py"), "avd_synth")
ref_video = _load(os.path.join(REF, "video.py"), "ref_video")
ref_fusion = _load(os.path.join(REF, "fusion.py"), "ref_fusion")
ref_hints = _load(os.path.join(REF, "heuristics_v2.py"), "ref_hints")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


class FakeCapture:
    def __init__(self, clip, fps):
        self.clip, self.fps, self.i = clip, fps, -1

    def isOpened(self):
        return True

    def get(self, prop):
        if prop == cv2.CAP_PROP_FPS: return float(self.fps)
        if prop == cv2.CAP_PROP_FRAME_WIDTH: return float(self.clip[0].shape[1])
        if prop == cv2.CAP_PROP_FRAME_HEIGHT: return float(self.clip[0].shape[0])
        if prop == cv2.CAP_PROP_FRAME_COUNT: return float(len(self.clip))
        return 0.0

    def grab(self):
        self.i += 1
        return self.i < len(self.clip)

    def retrieve(self):
        return True, self.clip[self.i]

    def release(self):
        pass


class Cv2Proxy:
    """cv2 with VideoCapture served from memory and a log of the analyzer's library results."""

    def __init__(self):
        self.sources, self.log = {}, {"area32": [], "thumb": [], "lap_var": [], "lap_sum": [], "lap_sumsq": [], "flow": []}

    def __getattr__(self, k):
        return getattr(cv2, k)

    def VideoCapture(self, path):
        if path in self.sources:
            return FakeCapture(*self.sources[path])
        return cv2.VideoCapture(path)

    def resize(self, src, dsize, *a, **kw):
        r = cv2.resize(src, dsize, *a, **kw)
        (self.log["area32"] if dsize == (32, 32) else self.log["thumb"]).append(r.copy())
        return r

    def Laplacian(self, src, ddepth, *a, **kw):
        r = cv2.Laplacian(src, ddepth, *a, **kw)
        self.log["lap_var"].append(float(r.var()))
        self.log["lap_sum"].append(int(r.sum()))
        self.log["lap_sumsq"].append(int((r * r).sum()))
        return r

    def calcOpticalFlowFarneback(self, *a, **kw):
        r = cv2.calcOpticalFlowFarneback(*a, **kw)
        self.log["flow"].append(r.copy())
        return r


def run_reference(clip_frames, fps, meta):
    proxy = Cv2Proxy()
    proxy.sources["mem://clip"] = (clip_frames, fps)
    ref_video.cv2 = proxy
    try:
        out = ref_video.analyze("mem://clip", dict(meta))
    finally:
        ref_video.cv2 = cv2
    return out, proxy.log


def pack_hash(area32):
    g = area32
    bits = (g >= g.mean()).astype(np.uint8).flatten()  # video.py:7-8 verbatim semantics
    return np.packbits(bits, bitorder="little").view(np.uint32)


def clip_case(name, n, H, W, seed, fps, meta_mode, arrays, keep_thumbs=(), keep_flows=()):
    clip = synth.Clip(n, H, W, seed, fps=fps)
    frames = [clip[i] for i in range(n)]
    meta = {"fps": fps, "width": W, "height": H, "duration": n / fps} if meta_mode == "full" else {}
    out, log = run_reference(frames, fps, meta)
    step = max(1, int(round((fps or 30) / 2)))
    sampled = frames[::step]
    hashes = [pack_hash(a) for a in log["area32"]]
    hams = [int(np.unpackbits((hashes[i] ^ hashes[i - 1]).view(np.uint8)).sum()) for i in range(1, len(hashes))]
    fstats = []
    for fl in log["flow"]:
        mag = np.sqrt(fl[..., 0] ** 2 + fl[..., 1] ** 2)
        fstats.append([float(np.mean(mag)), float(np.var(mag))])
    case = {
        "n_frames": n, "H": H, "W": W, "seed": seed, "fps": fps, "meta": meta, "step": step,
        "n_samples": len(sampled), "input_sha": sha(np.stack(sampled)),
        "result": out, "same_list_object": out["timeline"] is out["timeline_ai"],
        "area32_sha": [sha(a) for a in log["area32"]],
        "hash_words": [h.tolist() for h in hashes], "ham": hams,
        "thumb_sha": [sha(t) for t in log["thumb"]],
        "lap_var": log["lap_var"], "lap_sum": log["lap_sum"], "lap_sumsq": log["lap_sumsq"],
        "flow_stats": fstats,
    }
    arrays[f"{name}.area32"] = np.stack(log["area32"])
    for k in keep_thumbs:
        arrays[f"{name}.thumb{k}"] = log["thumb"][k]
    for k in keep_flows:
        arrays[f"{name}.flow{k}"] = log["flow"][k].astype(np.float32)
    return case, frames


def frame_case(H, W, seed, kind, arrays, keep_thumb=False):
    if kind == "noise":
        f = synth.noise_frames(1, H, W, seed)[0]
    else:
        f = synth.Clip(4, H, W, seed, events=False)[1]
    g = cv2.cvtColor(f, cv2.COLOR_BGR2GRAY)
    a = cv2.resize(g, (32, 32), interpolation=cv2.INTER_AREA)
    t = cv2.resize(g, (320, 320))
    lap = cv2.Laplacian(g, cv2.CV_64F)
    name = f"frame_{H}x{W}_{kind}_{seed}"
    arrays[name + ".area32"] = a
    if keep_thumb:
        arrays[name + ".thumb"] = t
    hsh = ref_video._average_hash(f, 32)  # the reference's own function
    assert np.array_equal(np.unpackbits(pack_hash(a).view(np.uint8), bitorder="little"), hsh)
    return {"name": name, "H": H, "W": W, "seed": seed, "kind": kind, "input_sha": sha(f), "gray_sha": sha(g),
            "hash_words": pack_hash(a).tolist(), "thumb_sha": sha(t), "lap_var": float(lap.var()),
            "lap_sum": int(lap.sum()), "lap_sumsq": int((lap * lap).sum())}


def fuse_cases():
    rng = np.random.default_rng(5)
    cases = []
    comps = [{"width": 1280, "height": 720, "fps": 30.0, "bit_rate": br} for br in (0, 900_000, 2_000_000, 3_500_000, 9_000_000)]
    for i in range(40):
        La, Lv = int(rng.integers(0, 12)), int(rng.integers(0, 12))
        audio = {"timeline": rng.uniform(0, 1, La).round(6).tolist(),
                 "flags_audio": {"speech_ratio": float(rng.choice([0.0, 0.2, 0.3, 0.9])), "tts_like": float(rng.choice([0.0, 0.5, 0.7, 0.97]))}}
        if i % 7 == 0:
            audio = {"timeline": [0.5] * La, "flags_audio": {"error": "x"}}
        lo, hi = [(0.0, 1.0), (0.0, 0.3), (0.7, 1.0), (0.4, 0.6)][i % 4]
        tl = rng.uniform(lo, hi, Lv).round(6).tolist()
        video = {"timeline": tl, "timeline_ai": tl,
                 "summary": {"dup_density": float(rng.choice([0.0, 0.01, 0.1, 0.3])), "scene_change_rate": float(rng.choice([0.0, 0.5, 0.75, 0.95])),
                             "flow_mean": float(rng.choice([0.1, 3.0, 6.0, 9.0])), "flow_var": 0.1,
                             "texture_var": float(rng.choice([10.0, 250.0, 400.0])), "w": 1280, "h": 720, "fps": 30.0}}
        if i % 11 == 0:
            video = {"timeline": [], "summary": {}, "timeline_ai": []}
        meta = comps[i % len(comps)]
        hints = ref_hints.compute_hints(dict(meta), "x")
        import copy
        a_in, v_in = copy.deepcopy(audio), copy.deepcopy(video)
        out = ref_fusion.fuse(audio, video, hints)
        # pre-rounding score (recomputed with the reference's own expression order is not exposed; store the
        # padded timelines so the test can check in-place extension, fusion.py:20-21)
        cases.append({"audio": a_in, "video": v_in, "meta": meta, "hints": hints, "out": out,
                      "audio_timeline_after": audio["timeline"], "video_timeline_after": video["timeline"]})
    return cases


def main():
    arrays = {}
    G = {"versions": {"cv2": cv2.__version__, "numpy": np.__version__, "python": sys.version.split()[0]},
         "reference_pins": "opencv-python-headless>=4.10,<4.11; numpy>=1.26,<2.0 (requirements.txt:7-8)",
         "clips": {}, "frames": [], "fuse": []}

    # --- clips through the unmodified reference analyze() -------------------------------------
    c, _ = clip_case("clip_small", 150, 270, 480, 11, 30.0, "full", arrays, keep_thumbs=(0, 5), keep_flows=(0, 4))
    G["clips"]["clip_small"] = c
    c, _ = clip_case("clip_odd", 110, 480, 854, 12, 25.0, "full", arrays, keep_thumbs=(1,))
    G["clips"]["clip_odd"] = c
    c, _ = clip_case("clip_nometa", 95, 240, 320, 13, 24.0, "empty", arrays)
    G["clips"]["clip_nometa"] = c
    c, frames = clip_case("clip_720p", 300, 720, 1280, 0, 30.0, "full", arrays, keep_flows=(6,))   # BASELINE configs[0]
    G["clips"]["clip_720p"] = c
    # the same clip through a real lossless file + the real cv2.VideoCapture (decode seam)
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "c.avi")
        vw = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"FFV1"), 30.0, (1280, 720))
        for f in frames:
            vw.write(f)
        vw.release()
        cap = cv2.VideoCapture(path)
        ok, f0 = cap.read()
        cap.release()
        assert ok and np.array_equal(f0, frames[0]), "FFV1 round trip is not lossless here"
        out_file = ref_video.analyze(path, {"fps": 30.0, "width": 1280, "height": 720, "duration": 10.0})
        out_probe = ref_video.analyze(path, {})
    G["clips"]["clip_720p"]["result_via_ffv1_file"] = out_file
    G["clips"]["clip_720p"]["result_via_ffv1_file_nometa"] = out_probe
    assert out_file == c["result"], "file path and in-memory path disagree"
    del frames
    # 1080p slice of BASELINE configs[1]: the first 12 samples (180 frames) of the seed-1 clip
    clip = synth.Clip(1800, 1080, 1920, 1, fps=30.0)
    fr = [clip[i] for i in range(0, 180)]
    out, log = run_reference(fr, 30.0, {"fps": 30.0, "width": 1920, "height": 1080, "duration": 6.0})
    fst = []
    for fl in log["flow"]:
        mag = np.sqrt(fl[..., 0] ** 2 + fl[..., 1] ** 2)
        fst.append([float(np.mean(mag)), float(np.var(mag))])
    G["clips"]["clip_1080p_head"] = {
        "n_frames": 180, "of_clip_frames": 1800, "H": 1080, "W": 1920, "seed": 1, "fps": 30.0, "step": 15, "n_samples": 12,
        "meta": {"fps": 30.0, "width": 1920, "height": 1080, "duration": 6.0}, "input_sha": sha(np.stack(fr[::15])),
        "result": out, "hash_words": [pack_hash(a).tolist() for a in log["area32"]], "thumb_sha": [sha(t) for t in log["thumb"]],
        "lap_var": log["lap_var"], "lap_sum": log["lap_sum"], "lap_sumsq": log["lap_sumsq"], "flow_stats": fst}
    del fr

    # --- single frames: every library code path -----------------------------------------------
    for (H, W, kind, keep) in [(1080, 1920, "noise", False), (1080, 1920, "tex", True), (2160, 3840, "tex", False),
                               (720, 1280, "noise", False), (480, 640, "tex", False),      # integer-scale INTER_AREA path
                               (960, 1280, "noise", False),                                  # integer-scale, 40x30
                               (640, 640, "tex", True),                                      # 2x2 shortcut of INTER_LINEAR
                               (480, 854, "noise", False), (333, 777, "noise", False),
                               (1440, 2560, "tex", False), (144, 256, "noise", True),        # thumbnail upsampling
                               (32, 32, "noise", False), (1920, 1080, "tex", False)]:
        G["frames"].append(frame_case(H, W, 3, kind, arrays, keep))

    # --- farneback pairs ------------------------------------------------------------------------
    G["flow_pairs"] = []
    tex = synth.texture(320, 420, 21, 0)[..., 1]
    tex2 = synth.texture(320, 320, 22, 1)[..., 2]
    rng = np.random.default_rng(23)
    pairs = {
        "shift4": (tex[:, 0:320], tex[:, 4:324]), "shift23": (tex[:, 0:320], tex[:, 23:343]),
        "static": (tex[:, 0:320], tex[:, 0:320]), "cut": (tex[:, 0:320], tex2),
        "noise": (rng.integers(0, 256, (320, 320), dtype=np.uint8), rng.integers(0, 256, (320, 320), dtype=np.uint8)),
        "static_noise": (tex[:, 0:320], np.clip(tex[:, 0:320].astype(np.int16) + rng.integers(-3, 4, (320, 320)), 0, 255).astype(np.uint8)),
        "flat": (np.full((320, 320), 128, np.uint8), np.full((320, 320), 128, np.uint8)),
    }
    for name, (p, n) in pairs.items():
        p, n = np.ascontiguousarray(p), np.ascontiguousarray(n)
        fl = cv2.calcOpticalFlowFarneback(p, n, None, 0.5, 3, 15, 3, 5, 1.2, 0)
        mag = np.sqrt(fl[..., 0] ** 2 + fl[..., 1] ** 2)
        arrays[f"pair_{name}.prev"], arrays[f"pair_{name}.next"] = p, n
        if name in ("shift4", "cut"):
            arrays[f"pair_{name}.flow"] = fl
        G["flow_pairs"].append({"name": name, "mean": float(np.mean(mag)), "var": float(np.var(mag)),
                                "flow_abs_max": float(np.abs(fl).max())})
    # library self-disagreement on the static pair (justifies the absolute floor of the flow tolerance)
    p, n = pairs["static"]
    a = cv2.calcOpticalFlowFarneback(np.ascontiguousarray(p), np.ascontiguousarray(n), None, 0.5, 3, 15, 3, 5, 1.2, 0)
    cv2.setUseOptimized(False)
    b = cv2.calcOpticalFlowFarneback(np.ascontiguousarray(p), np.ascontiguousarray(n), None, 0.5, 3, 15, 3, 5, 1.2, 0)
    cv2.setUseOptimized(True)
    ma, mb = np.sqrt((a ** 2).sum(2)), np.sqrt((b ** 2).sum(2))
    G["cv2_self_disagreement_static"] = {"mean_opt": float(ma.mean()), "mean_noopt": float(mb.mean()),
                                         "max_abs_flow_diff": float(np.abs(a - b).max())}

    G["fuse"] = fuse_cases()

    with open(os.path.join(OUT, "golden.json"), "w") as f:
        json.dump(G, f, indent=1)
    np.savez_compressed(os.path.join(OUT, "golden_arrays.npz"), **arrays)
    print("wrote", os.path.getsize(os.path.join(OUT, "golden.json")), os.path.getsize(os.path.join(OUT, "golden_arrays.npz")))


if __name__ == "__main__":
    main()

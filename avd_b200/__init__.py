"""avd-b200: B200-native per-frame video analyzer (drop-in for app/analyzers/video.py of
backtato/ai-video-detector).  Import as ``avd_b200`` (see /avd_b200.py at the repo root).

    from avd_b200 import video
    video.analyze(path, meta)            # same signature / return as the reference (video.py:10)
    video.analyze_frames(frames, meta)   # frames-level seam used by parity tests and bench

The CUDA extension (csrc/libavd.so, sm_100a) is mandatory: there is no CPU fallback.
"""
__version__ = "0.1.0"

"""Requests per second of small clips from 8 pool threads, with and without service.RequestBatcher:
python scripts/prof_batcher.py [requests_per_thread]"""
import os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import avd_b200.synth as synth
from avd_b200 import service, video
from avd_b200.engine import Engine
R = int(sys.argv[1]) if len(sys.argv) > 1 else 20
eng = Engine(0)
NT = 8
clips = [synth.Clip(150, 720, 1280, seed=200 + i, fps=30.0) for i in range(NT)]
reqs = []
for c in clips:                                   # configs[3]-like requests: 5 s of 720p -> 10 sampled frames, pinned
    pin = eng.pinned((10, 720, 1280, 3)); pin[:] = c.sampled(15)
    reqs.append(([pin[j] for j in range(10)], {"fps": 30.0, "width": 1280, "height": 720, "duration": 5.0}))


def run(fn):
    gate = threading.Barrier(NT + 1)

    def work(i):
        gate.wait()
        for _ in range(R):
            fn(*reqs[i])
    ts = [threading.Thread(target=work, args=(i,)) for i in range(NT)]
    [t.start() for t in ts]
    gate.wait()
    t0 = time.perf_counter()
    [t.join() for t in ts]
    return NT * R / (time.perf_counter() - t0)


direct = lambda f, m: video.analyze_frames(f, m, engine=eng)
run(direct)
a = run(direct)
pooled = lambda f, m: video.analyze_frames(f, m)              # engine=None: a context leased from the device's pool per request
run(pooled)
ap = run(pooled)
from avd_b200.engine import engine_pool
print(f"8 threads, 10 x 720p frames per request: one shared context {a:.0f} requests/s; pool of {engine_pool(0).size} contexts "
      f"(AVD_ENGINES_PER_DEVICE) {ap:.0f} requests/s ({10*ap:.0f} frames/s)")
for wait in (0.5, 1.0, 2.0):
    b = service.RequestBatcher(engine=eng, max_wait_ms=wait)
    run(b.analyze_frames)
    r = run(b.analyze_frames)
    print(f"8 threads, 10 x 720p frames per request: direct {a:.0f} requests/s ({10*a:.0f} frames/s); "
          f"batched (wait {wait} ms) {r:.0f} requests/s ({10*r:.0f} frames/s), mean batch {np.mean(b.batches):.1f}")

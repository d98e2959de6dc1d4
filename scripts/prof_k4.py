"""Device time of the audio feature kernels (K4) for a clip of `seconds` of 16 kHz audio: python scripts/prof_k4.py [seconds]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import avd_b200.synth as synth
from avd_b200 import audio
from avd_b200.engine import Engine
sec = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
eng = Engine(0)
wav = synth.audio(8, int(sec * 16000), "speech")
audio.analyze_samples(wav, 16000, engine=eng)
eng.profile_read(); eng.profile(True)
t0 = time.perf_counter()
for _ in range(5):
    out = audio.analyze_samples(wav, 16000, engine=eng)
dt = (time.perf_counter() - t0) / 5
eng.profile(False)
k = eng.profile_read()
print(f"audio {sec:.0f} s ({len(wav)//8000} windows): host call {1e3*dt:.2f} ms; " +
      "; ".join(f"{n} {1e3*v['ms']/v['launches']:.0f} us" for n, v in k.items() if n.startswith("k4_")))

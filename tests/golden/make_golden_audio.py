"""Generate tests/golden/golden_audio.json by EXECUTING THE REFERENCE'S OWN audio analyzer (build container only).

    python tests/golden/make_golden_audio.py

/root/reference/app/analyzers/audio.py is loaded by file path with an empty stand-in for the `soundfile` module (not
installed here; only its decode helper uses it).  Its analyze(path, meta) is run unmodified; the one hook is the
module-level decode helper _extract_wav_16k (audio.py:7-20: ffmpeg + soundfile, both absent here), replaced by a
function that returns the seeded samples -- exactly what the helper returns to analyze(): (tmp_path, float32 mono, sr).
Inputs are regenerated from seeds by avd_b200/synth.py:audio (integer PCM); their sha256 is stored.
"""
import hashlib
import importlib.util
import json
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference/app/analyzers"
OUT = os.path.dirname(os.path.abspath(__file__))


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


sys.modules.setdefault("soundfile", types.ModuleType("soundfile"))
synth = _load(os.path.join(ROOT, "avd_b200", "synth.py"), "avd_synth")
ref_audio = _load(os.path.join(REF, "audio.py"), "ref_audio")

CASES = [  # name, seed, samples, kind, sr
    ("speech_12s", 1, 192000, "speech", 16000),
    ("speech_ragged", 2, 100123, "speech", 16000),        # last window 4123 samples
    ("noise_3s", 3, 48000, "noise", 16000),
    ("tone_5s", 4, 80000, "tone", 16000),
    ("quiet_4s", 5, 64000, "quiet", 16000),
    ("silence_2s", 6, 32000, "silence", 16000),
    ("short_one_window", 7, 5000, "speech", 16000),      # a single ragged window, tlen 1
    ("speech_60s", 8, 960000, "speech", 16000),           # BASELINE configs[1]'s clip length
    ("odd_prime_tail", 9, 8000 * 3 + 4099, "noise", 16000),  # ragged window of prime length
    ("sr_8k", 10, 40000, "speech", 8000),                 # a different sample rate: 4000-sample windows
]


def main():
    G = {"versions": {"numpy": np.__version__, "python": sys.version.split()[0]}, "cases": []}
    for name, seed, n, kind, sr in CASES:
        wav = synth.audio(seed, n, kind)
        ref_audio._extract_wav_16k = lambda path, _w=wav, _sr=sr: (None, _w, _sr)
        with np.errstate(all="ignore"):
            out = ref_audio.analyze("in-memory", {"duration": n / sr})
        assert "error" not in out["flags_audio"], out
        G["cases"].append({"name": name, "seed": seed, "n": n, "kind": kind, "sr": sr,
                           "input_sha": hashlib.sha256(wav.tobytes()).hexdigest(), "out": out})
        print(name, len(out["timeline"]), out["scores"])
    with open(os.path.join(OUT, "golden_audio.json"), "w") as f:
        json.dump(G, f, indent=1)


if __name__ == "__main__":
    main()

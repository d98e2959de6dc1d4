"""sha256 of the dense flow fields of a fixed set of thumbnail pairs (to check that a K2 change is bit-neutral)."""
import hashlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import avd_b200.synth as synth
from avd_b200.engine import Engine
eng = Engine(0)
n = 24
tex = synth.texture(320, 320 + 4 * n, 5, 0)[..., 1]
thumbs = np.stack([tex[:, 3 * i:3 * i + 320] for i in range(n)])
thumbs[7] = np.random.default_rng(1).integers(0, 256, (320, 320), dtype=np.uint8)      # a cut
thumbs[12] = thumbs[11]                                                                  # a static pair
t = torch.from_numpy(thumbs).cuda()
f = eng.flow_fields(t[:-1].contiguous(), t[1:].contiguous()).cpu().numpy()
print(os.environ.get("AVD_LIB", "default"), hashlib.sha256(f.tobytes()).hexdigest(), float(np.abs(f).mean()))

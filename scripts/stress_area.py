"""Randomised stress of the INTER_AREA certificate + fix-up: streaming kernel (and, with `rows` as argv[1], the any-width
kernel on ragged / fractional-scale shapes) vs the generic kernel (library order) on frames built to put many cell means next
to rounding boundaries.  Prints mismatch counts (must be 0)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from avd_b200.engine import Engine
eng = Engine(0)
g = torch.Generator(device="cuda"); g.manual_seed(7)
bad = tot = flagged_like = 0
ROWS = len(sys.argv) > 1 and sys.argv[1] == "rows"
MODE = 3 if ROWS else 0
SHAPES = [(480, 854), (768, 1366), (333, 777), (540, 960), (1080, 1918), (33, 35), (600, 1920), (97, 4001)] if ROWS else \
         [(720, 1280), (1080, 1920), (2160, 3840), (648, 1344), (1000, 2048)]
for (H, W) in SHAPES:
    for kind in range(8):
        n = (96 if H * W < 600000 else 24) if H < 2000 else 6
        base = torch.randint(0, 255, (n, 1, 1, 1), device="cuda", generator=g).float()
        if kind == 0:   # two adjacent levels, random mix per frame: means spread over [v, v+1]
            pr = torch.rand((n, 1, 1, 1), device="cuda", generator=g)
            fr = base + (torch.rand((n, H, W, 1), device="cuda", generator=g) < pr).float()
        elif kind == 1: # exactly half/half columns
            fr = base + (torch.arange(W, device="cuda") % 2).view(1, 1, W, 1).float().expand(n, H, W, 1)
        elif kind == 2: # smooth gradient + small noise
            fr = base * 0 + torch.linspace(0, 255, W, device="cuda").view(1, 1, W, 1) * torch.rand((n, 1, 1, 1), device="cuda", generator=g) \
                 + torch.randint(0, 3, (n, H, W, 1), device="cuda", generator=g).float()
        elif kind == 3: # sparse impulses on a constant
            fr = base + 200 * (torch.rand((n, H, W, 1), device="cuda", generator=g) < 0.0025).float()
        elif kind == 4: # row stripes (fractional row weights matter)
            fr = base + (torch.arange(H, device="cuda") % 2).view(1, H, 1, 1).float().expand(n, H, W, 1)
        elif kind == 5: # quarter mix
            fr = base + (torch.rand((n, H, W, 1), device="cuda", generator=g) < 0.25).float() * 2
        elif kind == 6: # full noise
            fr = torch.randint(0, 256, (n, H, W, 1), device="cuda", generator=g).float()
        else:           # blocks of 3 levels
            fr = base + torch.randint(0, 3, (n, H // 8 + 1, W // 8 + 1, 1), device="cuda", generator=g).float() \
                .repeat_interleave(8, 1).repeat_interleave(8, 2)[:, :H, :W]
        fr = fr.clamp(0, 255).to(torch.uint8).expand(n, H, W, 3).contiguous()     # B=G=R=v -> gray = v
        outs = []
        for mode in (MODE, 1):
            eng.set_k1_mode(mode)
            rec, th, ar = eng.frames_features(fr, want_area32=True)
            outs.append((ar.clone(), th.clone(), rec.clone()))
        eng.set_k1_mode(0)
        d = int((outs[0][0] != outs[1][0]).sum()); dt = int((outs[0][1] != outs[1][1]).sum()); dr = int((outs[0][2] != outs[1][2]).sum())
        bad += d + dt + dr; tot += n * 1024
        print(f"{H}x{W} kind {kind}: area mismatches {d}, thumbnail {dt}, record bytes {dr}", flush=True)
print("TOTAL cells", tot, "mismatches", bad)
sys.exit(1 if bad else 0)

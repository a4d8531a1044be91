#!/usr/bin/env python
"""How much do finishing (re-spawning) worlds cost a single-wave launch?  Per-step time vs. number of done worlds."""
import sys
from pathlib import Path
import numpy as np
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from as_cops_and_thieves_b200.worlds import CatWorlds  # noqa: E402
import parity_utils as pu  # noqa: E402
N = 4096
cw = CatWorlds(pu.named_cmap("squarinth"), N, want_f32=False, want_shared=False, seed=0)
cw.reset()
g = torch.Generator(device="cuda").manual_seed(1)
acts = [torch.randint(0, 4, (N, 3), dtype=torch.uint8, device="cuda", generator=g) for _ in range(16)]
K = 2400
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
done = torch.zeros(K, dtype=torch.int32, device="cuda")
for i in range(K):
    ev[i][0].record(); cw.step(acts[i % 16]); ev[i][1].record()
    done[i] = cw.terminated.sum()
torch.cuda.synchronize()
t = np.array([a.elapsed_time(b) * 1e3 for a, b in ev]); d = done.cpu().numpy()
for lo, hi in ((0, 0), (1, 5), (6, 15), (16, 40), (41, 200), (201, 5000)):
    m = (d >= lo) & (d <= hi)
    if m.any():
        print(f"done worlds {lo:4d}..{hi:4d}: {m.sum():5d} steps, median {np.median(t[m]):6.1f} us, mean {t[m].mean():6.1f} us")
print(f"all: mean {t.mean():.1f} us, median {np.median(t):.1f} us, mean done/step {d.mean():.1f}")

#!/usr/bin/env python
"""Long-run invariants (GPU box): many thousands of steps per map; state stays finite, speeds stay clamped,
agents stay near the map, counters behave.  A robustness soak, not a parity test."""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from as_cops_and_thieves_b200.worlds import CatWorlds  # noqa: E402
import parity_utils as pu  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
for mp, free, N in (("squarinth", False, 2048), ("lbirinth", False, 2048), ("grandbyrinth", False, 2048),
                    ("labyrinth", True, 2048), ("agh-map", True, 2048), ("agh-map", False, 1024)):
    cmap = pu.named_cmap(mp, free_spawn=free)
    cw = CatWorlds(cmap, N, want_f32=False, want_shared=True, seed=4)
    cw.reset()
    g = torch.Generator(device="cuda").manual_seed(5)
    acts = [torch.randint(0, 4, (N, cw.A), dtype=torch.uint8, device="cuda", generator=g) for _ in range(64)]
    lo = torch.tensor([cmap.grid_x0, cmap.grid_y0], device="cuda") - 3000
    hi = torch.tensor([cmap.grid_x0 + cmap.nx * cmap.cell, cmap.grid_y0 + cmap.ny * cmap.cell], device="cuda") + 3000
    done_total = cop_wins = 0
    for i in range(steps):
        cw.step(acts[i % 64])
        if i % 500 == 499 or i == steps - 1:
            st = cw.get_state()
            pos, vel, vb = st["pos"], st["vel"], st["vbias"]
            assert torch.isfinite(pos).all() and torch.isfinite(vel).all() and torch.isfinite(vb).all(), (mp, i)
            # the clamp to 125 is applied with the action (entity.py:133-134); contact impulses of the step that follows
            # can push a body above it until its next action, exactly as in the reference
            assert (vel.norm(dim=-1) <= 250.0).all(), (mp, i, float(vel.norm(dim=-1).max()))
            assert ((pos >= lo) & (pos <= hi)).all(), (mp, i)
            assert (st["step_count"] >= 0).all() and (st["step_count"] < 400).all(), (mp, i)
            d = cw.obs_dist.float()
            # alpha = 0 hits report the ray END through the float16 chain (SURVEY C-5): up to ~1 above the 400 range
            assert torch.isfinite(d).all() and (d >= 0).all() and (d <= 402.0).all(), (mp, i, float(d.max()))
            assert ((cw.obs_type <= 2) | (cw.obs_type == 4)).all(), (mp, i)
            assert torch.isfinite(cw.reward).all() and (cw.reward.abs() <= 1.5).all()
        done_total += int(cw.terminated.sum()) if i % 50 == 0 else 0
    st = cw.get_state()
    print(f"{mp}{'-free' if free else ''}: {steps} steps x {N} worlds ok; episodes per world {float(st['episode'].float().mean()):.1f}, "
          f"max |pos| {float(st['pos'].abs().max()):.0f}", flush=True)
    cw.close()

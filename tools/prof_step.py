#!/usr/bin/env python
"""Tiny driver for ncu: a few env steps of one map so a single launch can be captured."""
import argparse
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from as_cops_and_thieves_b200.worlds import CatWorlds  # noqa: E402
import parity_utils as pu  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--map", default="squarinth")
ap.add_argument("--worlds", type=int, default=4096)
ap.add_argument("--steps", type=int, default=60)
ap.add_argument("--free", type=int, default=0)
ap.add_argument("--cell", type=float, default=None)
ap.add_argument("--flush", default="none", choices=["none", "read", "write"])
ap.add_argument("--ray-cell", type=float, default=0.0, help="ray_list_cell: 0 auto, < 0 rasterise")
a = ap.parse_args()
kw = {} if a.cell is None else {"cell": a.cell}
cmap = pu.named_cmap(a.map, free_spawn=bool(a.free), **kw)
cw = CatWorlds(cmap, a.worlds, want_f32=False, want_shared=False, ray_list_cell=a.ray_cell)
cw.reset()
acts = [torch.randint(0, 4, (a.worlds, cw.A), dtype=torch.uint8, device="cuda") for _ in range(8)]
for i in range(a.steps):
    cw.step(acts[i % 8])
torch.cuda.synchronize()
if a.flush == "none":
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(50):
        cw.step(acts[i % 8])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 50
else:
    flush = torch.zeros(192 * 1024 * 1024 // 4, dtype=torch.int32, device="cuda")
    ev = []
    for i in range(50):
        flush.add_(1) if a.flush == "write" else flush.sum()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); cw.step(acts[i % 8]); e1.record()
        ev.append((e0, e1))
    torch.cuda.synchronize()
    ts = sorted(x.elapsed_time(y) for x, y in ev)
    ms = sum(ts) / len(ts)
    print(f"flush={a.flush}: min {ts[0]*1e3:.1f} median {ts[25]*1e3:.1f} max {ts[-1]*1e3:.1f} us")
print(f"ray lists: {cw.info.ray_list_nx}x{cw.info.ray_list_ny} cells of {cw.info.ray_list_cell:.1f}, {cw.info.ray_list_bytes/1e6:.1f} MB; overflow counts {cw.overflow_counts()}")
print(f"{a.map} N={a.worlds} cell={cmap.cell:.1f}: {ms*1e3:.1f} us/step {a.worlds*cw.A/ms*1e3:.3e} agent-steps/s grid {cw.info.grid} x {cw.info.warps_per_cta} warps")

#!/usr/bin/env python
"""Host-side model of the per-(cell, ray) candidate lists: list sizes and the number of (edge, ray) tests a
lane-per-ray walk with near-to-far early exit performs, per ray and as the maximum over each 32-ray group
(what a warp pays).  usage: ray_list_stats.py <map> [cell] [n_origins]"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from as_cops_and_thieves_b200.maps import ray_lists  # noqa: E402
import parity_utils as pu  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "agh-map"
cell = float(sys.argv[2]) if len(sys.argv) > 2 else None
n_org = int(sys.argv[3]) if len(sys.argv) > 3 else 600
free = name in ("agh-map", "labyrinth")
cmap = pu.named_cmap(name, free_spawn=free, **({"cell": cell} if cell else {}))
R, L, rsum = 90, 400.0, 2.0
import time
t0 = time.time()
off, ent = ray_lists(cmap, R, L, rsum)
print(f"{name}: cell {cmap.cell:.1f} grid {cmap.nx}x{cmap.ny} E={cmap.n_edges}  build {time.time() - t0:.1f}s  "
      f"entries {len(ent)} ({len(ent) * 4 / 1e6:.2f} MB + {len(off) * 4 / 1e6:.2f} MB offsets)  mean list {len(ent) / (len(off) - 1):.1f}  "
      f"max {np.diff(off).max()}")

# sample origins = spawn positions of the oracle (free space)
from oracle.cat_oracle import Oracle  # noqa: E402
orc = Oracle(cmap, seed=0)
st = orc.new_state(n_org)
orc.reset(st)
org = st.pos.reshape(-1, 2)[:n_org]
E = cmap.n_edges
prev = np.zeros(E, np.int64)
for h in range(cmap.n_hulls):
    o, e = int(cmap.hull_off[h]), int(cmap.hull_off[h + 1])
    prev[o:e] = np.roll(np.arange(o, e), 1)
B, n, ln = cmap.vert, cmap.normal, cmap.edge_len
ang = np.arange(R) * 2 * np.pi / R
U = np.stack([np.cos(ang), np.sin(ang)], 1)
tests = np.zeros((len(org), R), np.int64)
for oi, o in enumerate(org):
    cx, cy = int((o[0] - cmap.grid_x0) / cmap.cell), int((o[1] - cmap.grid_y0) / cmap.cell)
    c = cy * cmap.nx + cx
    for i in range(R):
        q0, q1 = off[c * R + i], off[c * R + i + 1]
        best = L
        k = 0
        for q in range(q0, q1):
            e_, lb = int(ent[q] & 0xFFFF), (int(ent[q]) >> 16) / 64.0
            if best < lb:
                break
            k += 1
            r = o - B[e_]
            d = r @ n[e_] - rsum
            un = U[i] @ n[e_]
            s = np.inf
            if d >= 0 and un < 0:
                sp = d / -un
                cr = n[e_][0] * (sp * U[i][1] + r[1]) - n[e_][1] * (sp * U[i][0] + r[0])
                if -ln[e_] <= cr <= 0:
                    s = sp
            cp = r[0] * U[i][1] - r[1] * U[i][0]
            disc = rsum * rsum - cp * cp
            if disc >= 0:
                sc = -(r @ U[i]) - np.sqrt(disc)
                if 0 <= sc < s:
                    s = sc
            best = min(best, s)
        tests[oi, i] = k
grp = [tests[:, 0:32], tests[:, 32:64], tests[:, 64:90]]
mx = np.stack([g.max(1) for g in grp], 1)
print(f"tests per ray: mean {tests.mean():.2f}  p50 {np.median(tests):.0f}  p95 {np.percentile(tests, 95):.0f}  max {tests.max()}")
print(f"max over a 32-ray group: mean {mx.mean():.2f}  p95 {np.percentile(mx, 95):.0f};  per agent sum of group maxima {mx.sum(1).mean():.1f}")
print(f"lane efficiency of the walk: {tests.sum() / (mx[:, 0] * 32 + mx[:, 1] * 32 + mx[:, 2] * 26).sum():.2f}")

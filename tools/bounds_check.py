#!/usr/bin/env python
"""Checked build of the kernels (stand-in for compute-sanitizer, which is closed on this GPU pool): -DCAT_STATS turns every
CAT_CHECK in csrc/world_kernel.cuh into a counted index / capacity assertion (shared-memory staging areas, candidate
queue, contact slots, ray-list slot / overflow offsets, edge ids).  Runs every map through both sensor paths, across
episode ends, plus ragged configurations, and prints the violation count (must be 0).
usage: bounds_check.py [--build]"""
import ctypes as C
import os
import subprocess
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from as_cops_and_thieves_b200 import build  # noqa: E402
out = ROOT / "as_cops_and_thieves_b200" / "variants" / "libcat_stats.so"
out.parent.mkdir(exist_ok=True)
if "--build" in sys.argv or not out.exists():
    r = subprocess.run(build.nvcc_cmd(out, ["-DCAT_STATS"]), capture_output=True, text=True)
    print("built", out, r.returncode)
    if "--build" in sys.argv:
        sys.exit(r.returncode)
os.environ["CAT_B200_LIB"] = str(out)
import torch  # noqa: E402
from as_cops_and_thieves_b200 import _lib  # noqa: E402
from as_cops_and_thieves_b200.maps import Map, compile_map  # noqa: E402
from as_cops_and_thieves_b200.worlds import CatWorlds  # noqa: E402
import parity_utils as pu  # noqa: E402
import json, tempfile  # noqa: E402
sys.path.insert(0, str(ROOT / "tests"))
from test_parity_gpu import RAGGED_MAP  # noqa: E402
L = _lib.load()
buf = (C.c_ulonglong * 8)()
L.cat_debug_stats(buf, 1)
total = 0
cases = [(n, f, {}) for n, f in (("squarinth", False), ("lbirinth", False), ("grandbyrinth", False), ("labyrinth", True),
                                 ("agh-map", True), ("agh-map", False))]
with tempfile.NamedTemporaryFile("w", suffix=".json", delete=False) as f:
    json.dump(RAGGED_MAP, f)
ragged = compile_map(Map(f.name), name="ragged")
for name, free, kw in cases + [("ragged", None, dict(n_rays=45)), ("ragged", None, dict(n_rays=128, dt=1 / 15))]:
    cmap = ragged if name == "ragged" else pu.named_cmap(name, free_spawn=free)
    for ray_cell in (0.0, -1.0):
        N = 2048
        cw = CatWorlds(cmap, N, want_f32=True, want_shared=True, want_critic=True, want_hits=True, max_step_count=25,
                       ray_list_cell=ray_cell, **kw)
        cw.reset()
        g = torch.Generator().manual_seed(0)
        for i in range(60):                     # crosses two episode ends: auto-reset, second staging of the slots
            cw.step(torch.randint(0, 4, (N, cw.A), dtype=torch.uint8, generator=g).cuda())
        cw.step_host(torch.randint(0, 4, (N, cw.A), dtype=torch.uint8, generator=g), mode="pipelined", chunks=3)
        cw.step_host(torch.randint(0, 4, (N, cw.A), dtype=torch.uint8, generator=g), mode="zero_copy")
        cw.step_host(torch.randint(0, 4, (N, cw.A), dtype=torch.uint8, generator=g), mode="zero_copy", packed=False)
        cw.step_host(torch.randint(0, 4, (N, cw.A), dtype=torch.uint8, generator=g), mode="staged", packed=False)
        torch.cuda.synchronize()
        L.cat_debug_stats(buf, 1)
        print(f"{name:12s} free={free} {kw} sensor={'lists' if ray_cell == 0 else 'rasteriser'}: {buf[6]} violations"
              + (f" (last at world_kernel.cuh:{buf[7]})" if buf[6] else "") + f", capacity overflows {cw.overflow_counts()}")
        total += buf[6]
        cw.close()
print("TOTAL violations:", total)
sys.exit(1 if total else 0)

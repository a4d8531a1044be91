#!/usr/bin/env python
"""Rasteriser work counters from a -DCAT_STATS build (developer tool): candidates, (edge, ray) pairs, occlusion culls."""
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
if "--build" in sys.argv:
    r = subprocess.run(build.nvcc_cmd(out, ["-DCAT_STATS"]), capture_output=True, text=True)
    print("built", out, r.returncode)
    sys.exit(r.returncode)
os.environ["CAT_B200_LIB"] = str(out)
import torch  # noqa: E402
from as_cops_and_thieves_b200 import _lib  # noqa: E402
from as_cops_and_thieves_b200.worlds import CatWorlds  # noqa: E402
import parity_utils as pu  # noqa: E402
L = _lib.load()
for mp, free, N in (("squarinth", False, 4096), ("lbirinth", False, 4096), ("grandbyrinth", False, 4096), ("labyrinth", True, 4096), ("agh-map", True, 4096)):
    cw = CatWorlds(pu.named_cmap(mp, free_spawn=free), N, want_f32=False, want_shared=False)
    cw.reset()
    acts = [torch.randint(0, 4, (N, 3), dtype=torch.uint8, device="cuda") for _ in range(8)]
    for i in range(100):
        cw.step(acts[i % 8])
    buf = (C.c_ulonglong * 8)()
    L.cat_debug_stats(buf, 1)
    K = 50
    for i in range(K):
        cw.step(acts[i % 8])
    L.cat_debug_stats(buf, 1)
    sweeps = N * 3 * K
    print(f"{mp}: per agent sweep: candidates {buf[0]/sweeps:.1f}, pairs {buf[1]/sweeps:.1f}, narrow (<=3 rays) {buf[2]/sweeps:.1f}, "
          f"occlusion-culled {buf[3]/sweeps:.1f}; list walk: {buf[4]/sweeps:.1f} edge tests per sweep ({buf[4]/sweeps/cw.R:.2f} per ray), "
          f"{buf[5]/(N*K):.1f} warp iterations per world-step, lane efficiency {buf[4]/max(buf[5],1)/32:.2f}")
    cw.close()

#!/usr/bin/env python
"""Time CatWorlds.step_host(zero_copy=True) for several pinned-buffer layouts (GPU box)."""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from as_cops_and_thieves_b200.worlds import CatWorlds  # noqa: E402
import parity_utils as pu  # noqa: E402

for mp, free, N in (("squarinth", False, 4096), ("agh-map", True, 16384)):
    cmap = pu.named_cmap(mp, free_spawn=free)
    for layout in ("record128", "staged", "pipe1", "pipe2", "pipe3", "pipe4", "pipe6", "pipe8"):
        cw = CatWorlds(cmap, N, want_f32=False, want_shared=False)
        cw._zc_layout = layout
        cw.reset()
        acts = [torch.randint(0, 4, (N, cw.A), dtype=torch.uint8).pin_memory() for _ in range(8)]
        kw = dict(mode="staged") if layout == "staged" else (dict(mode="pipelined", chunks=int(layout[4:])) if layout.startswith("pipe") else dict(mode="zero_copy"))
        for i in range(10):
            cw.step_host(acts[i % 8], **kw)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        K = 300
        e0.record()
        for i in range(K):
            cw.step_host(acts[i % 8], **kw)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / K
        print(f"{mp} N={N} {layout:10s}: {ms*1e3:7.1f} us/step  {N*cw.A/ms*1e3:.3e} agent-steps/s  {cw.d2h_bytes_per_step/ms/1e6:.1f} GB/s", flush=True)
        cw.close()

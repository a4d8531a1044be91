#!/usr/bin/env python
"""Time CatWorlds.step_host for every transfer strategy / chunk count (GPU box).  usage: host_modes.py [map worlds free]"""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from as_cops_and_thieves_b200.worlds import CatWorlds  # noqa: E402
import parity_utils as pu  # noqa: E402
cases = [("squarinth", 4096, False), ("agh-map", 16384, True)] if len(sys.argv) < 4 else [(sys.argv[1], int(sys.argv[2]), bool(int(sys.argv[3])))]
for name, N, free in cases:
    cw = CatWorlds(pu.named_cmap(name, free_spawn=free), N, want_f32=False, want_shared=False)
    cw.reset()
    acts = [torch.randint(0, 4, (N, cw.A), dtype=torch.uint8).pin_memory() for _ in range(8)]
    dacts = [a.cuda() for a in acts]
    for i in range(50):
        cw.step(dacts[i % 8])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(200):
        cw.step(dacts[i % 8])
    e1.record(); torch.cuda.synchronize()
    print(f"{name} x {N}: device-resident {e0.elapsed_time(e1) / 200 * 1e3:.1f} us/step; record {cw.record_bytes} B/world, {N * cw.record_bytes / 1e6:.2f} MB/step")
    modes = [("zero_copy", None), ("staged", None)] + [("pipelined", c) for c in (1, 2, 3, 4, 6, 8, 12, 16)]
    modes = [(m, c, True) for m, c in modes] + [("zero_copy", None, False), ("pipelined", 4, False)]
    for mode, ch, packed in modes:
        for i in range(20):
            cw.step_host(acts[i % 8], mode=mode, chunks=ch, packed=packed)
        import time
        t0 = time.perf_counter()
        K = 300
        for i in range(K):
            cw.step_host(acts[i % 8], mode=mode, chunks=ch, packed=packed)
        dt = (time.perf_counter() - t0) / K
        print(f"  {mode:10s} chunks={ch} {'packed' if packed else 'u8    '}: {dt * 1e6:7.1f} us/step  {cw.d2h_bytes(packed) / dt / 1e9:5.1f} GB/s D2H")
    cw.close()

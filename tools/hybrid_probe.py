#!/usr/bin/env python
"""Feasibility probe: do DMA copies and zero-copy kernel stores add up on the PCIe link?  Half A of the worlds
steps into device buffers and is DMA-copied on a second stream while half B steps with zero-copy stores."""
import ctypes as C
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from as_cops_and_thieves_b200 import _lib  # noqa: E402
from as_cops_and_thieves_b200.worlds import CatWorlds  # noqa: E402
import parity_utils as pu  # noqa: E402

for mp, free, N in (("squarinth", False, 4096), ("agh-map", True, 16384)):
    cmap = pu.named_cmap(mp, free_spawn=free)
    for frac in (0.0, 0.3, 0.4, 0.5, 0.6, 0.7, 1.0):
        nA = int(N * frac) // 8 * 8
        nB = N - nA
        A = CatWorlds(cmap, nA, want_f32=False, want_shared=False) if nA else None
        B = CatWorlds(cmap, nB, gid0=nA, want_f32=False, want_shared=False) if nB else None
        s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
        for w in (A, B):
            if w is not None:
                w.reset()
        torch.cuda.synchronize()
        actsA = [torch.randint(0, 4, (max(nA, 1), 3), dtype=torch.uint8).pin_memory() for _ in range(4)]
        actsB = [torch.randint(0, 4, (max(nB, 1), 3), dtype=torch.uint8).pin_memory() for _ in range(4)]
        hA = A._host_buffers(False) if A else None
        hB = B._host_buffers(True) if B else None
        ioA = None
        if A:
            ioA = A._make_io()
        ev = torch.cuda.Event()

        def step(i):
            with torch.cuda.stream(s1):
                if A:
                    ioA.actions, ioA.actions_kind = actsA[i % 4].data_ptr(), 0
                    _lib.check(A.L.cat_env_step(A._h, A.state.data_ptr(), C.byref(ioA), s1.cuda_stream), "a")
                    ev.record(s1)
                if B:
                    io = hB["io"]
                    io.actions, io.actions_kind = actsB[i % 4].data_ptr(), 0
                    _lib.check(B.L.cat_env_step(B._h, B.state.data_ptr(), C.byref(io), s1.cuda_stream), "b")
            if A:
                with torch.cuda.stream(s2):
                    s2.wait_event(ev)
                    hA["blob"].copy_(A._out, non_blocking=True)
            s1.synchronize(); s2.synchronize()
        for i in range(10):
            step(i)
        import time
        t0 = time.perf_counter()
        K = 300
        for i in range(K):
            step(i)
        us = (time.perf_counter() - t0) / K * 1e6
        print(f"{mp} N={N} DMA share {frac:.1f}: {us:7.1f} us/step  {N*3/us*1e6:.3e} agent-steps/s", flush=True)
        for w in (A, B):
            if w is not None:
                w.close()

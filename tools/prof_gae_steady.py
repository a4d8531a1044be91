#!/usr/bin/env python
"""Steady-state throughput of the GAE + advantage-normalisation kernels: launches back to back over rotating buffer sets
(several times the L2), ONE CUDA-event pair around the lot — every byte, including the dirty lines a single launch leaves
in L2, has to reach HBM inside the timed span.  usage: prof_gae_steady.py [--T 256] [--M 49152] [--lib path]"""
import argparse, os, sys
from pathlib import Path
ap = argparse.ArgumentParser()
ap.add_argument("--T", type=int, default=256)
ap.add_argument("--M", type=int, default=49152)
ap.add_argument("--sets", type=int, default=0)
ap.add_argument("--rounds", type=int, default=6)
ap.add_argument("--lib", default="")
a = ap.parse_args()
if a.lib:
    os.environ["CAT_B200_LIB"] = a.lib
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from as_cops_and_thieves_b200 import _lib  # noqa: E402
L = _lib.load()
dev = torch.device("cuda:0")
T, M = a.T, a.M
n = T * M
S = a.sets or max(3, int(-(-4 * 126e6 // (25 * n))))
g = torch.Generator(device=dev).manual_seed(7)
sets = []
for _ in range(S):
    r = torch.randn((T, M), device=dev, generator=g); v = torch.randn((T, M), device=dev, generator=g)
    d = (torch.rand((T, M), device=dev, generator=g) < 0.01).to(torch.uint8)
    sets.append((r, v, d, torch.randn((M,), device=dev, generator=g), torch.empty_like(r), torch.empty_like(r),
                 torch.zeros(6, dtype=torch.float64, device=dev)))
stream = torch.cuda.current_stream(dev).cuda_stream


def gae(s):
    r, v, d, lv, ret, adv, st = s
    _lib.check(L.cat_gae(r.data_ptr(), d.data_ptr(), v.data_ptr(), lv.data_ptr(), ret.data_ptr(), adv.data_ptr(), st.data_ptr(),
                         T, M, 0.99, 0.95, stream), "gae")


def norm(s):
    _lib.check(L.cat_adv_normalize(s[5].data_ptr(), n, s[6].data_ptr(), n, stream), "norm")


def run(fns, bytes_per_sample, name):
    res = []
    for rep in range(5):
        for s in sets:
            for f in fns:
                f(s)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.rounds):
            for s in sets:
                for f in fns:
                    f(s)
        e1.record(); torch.cuda.synchronize()
        res.append(e0.elapsed_time(e1) / (a.rounds * S))
    res.sort()
    med = res[len(res) // 2]
    print(f"{name:28s} T={T} M={M} {S} sets: {med * 1e3:7.1f} us per launch(es)  {bytes_per_sample * n / med / 1e6:7.1f} GB/s "
          f"= {bytes_per_sample * n / med / 1e6 / 6541.1:.3f} of 6541 GB/s", flush=True)


run([gae], 17, "cat_gae steady state")
run([norm], 8, "cat_adv_normalize steady")
run([gae, norm], 25, "gae + normalize steady")

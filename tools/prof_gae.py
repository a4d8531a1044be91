#!/usr/bin/env python
"""GAE kernels vs. plain torch streaming kernels under the same timing method (GPU box)."""
import argparse
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from as_cops_and_thieves_b200 import _lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--T", type=int, default=256)
ap.add_argument("--M", type=int, default=49152)
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--flush", default="read", choices=["read", "write", "none"])
a = ap.parse_args()
L = _lib.load()
dev = torch.device("cuda:0")
T, M = a.T, a.M
g = torch.Generator(device=dev).manual_seed(7)
r = torch.randn((T, M), device=dev, generator=g)
v = torch.randn((T, M), device=dev, generator=g)
d = (torch.rand((T, M), device=dev, generator=g) < 0.01).to(torch.uint8)
lv = torch.randn((M,), device=dev, generator=g)
ret, adv = torch.empty_like(r), torch.empty_like(r)
stats = torch.zeros(6, dtype=torch.float64, device=dev)
flush = torch.zeros(192 * 1024 * 1024 // 4, dtype=torch.int32, device=dev)
stream = torch.cuda.current_stream(dev).cuda_stream
n = T * M


def do_flush():
    if a.flush == "read":
        flush.sum()
    elif a.flush == "write":
        flush.add_(1)


def timed(fn, bytes_per_sample, name):
    ms = []
    for i in range(-3, a.iters):
        do_flush()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        if i >= 0:
            ms.append(e0.elapsed_time(e1))
    ms.sort()
    med = ms[len(ms) // 2]
    print(f"{name:34s} T={T} M={M} flush={a.flush}: median {med*1e3:7.1f} us  min {ms[0]*1e3:7.1f} us  "
          f"{bytes_per_sample*n/med/1e6:7.1f} GB/s (median)", flush=True)


timed(lambda: _lib.check(L.cat_gae(r.data_ptr(), d.data_ptr(), v.data_ptr(), lv.data_ptr(), ret.data_ptr(), adv.data_ptr(),
                                   stats.data_ptr(), T, M, 0.99, 0.95, stream), "gae"), 17, "cat_gae (17 B/sample)")
timed(lambda: _lib.check(L.cat_adv_normalize(adv.data_ptr(), n, stats.data_ptr(), n, stream), "norm"), 8, "cat_adv_normalize (8 B/sample)")
timed(lambda: adv.mul_(1.0001), 8, "torch adv.mul_ in place (8 B)")
timed(lambda: ret.copy_(r), 8, "torch ret.copy_(r) (8 B)")
timed(lambda: torch.add(r, v, out=ret), 12, "torch add(r, v, out=ret) (12 B)")

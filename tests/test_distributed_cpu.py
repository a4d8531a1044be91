"""N>1 host path on CPU: world_size-2 gloo process group."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parents[1]


def _worker(rank: int, world_size: int, port: int, tmp: str):
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world_size),
                      LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        from as_cops_and_thieves_b200.sharding import allreduce_gradients, dist_env, shard_range
        import parity_utils as pu
        from oracle.cat_oracle import Oracle

        assert dist_env() == (rank, rank, world_size)
        # (1) gradient all-reduce: flat bucket, mean over ranks
        torch.manual_seed(rank)
        net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3))
        for p in net.parameters():
            p.grad = torch.full_like(p, float(rank + 1))
        flat = allreduce_gradients(net.parameters(), world_size)
        expect = sum(range(1, world_size + 1)) / world_size
        assert flat.numel() == sum(p.numel() for p in net.parameters())
        for p in net.parameters():
            assert torch.allclose(p.grad, torch.full_like(p, expect))
        # (2) world sharding: each rank resets its contiguous shard with global ids; the union equals
        #     one unsharded reset (spawn RNG is keyed by global world id, not by rank)
        n_global = 37
        gid0, n_local = shard_range(n_global, rank, world_size)
        cm = pu.named_cmap("squarinth")
        orc = Oracle(cm, seed=21)
        st = orc.new_state(n_local, gid0=gid0)
        orc.reset(st)
        mine = torch.zeros(n_global, 3, 2, dtype=torch.float64)
        mine[gid0:gid0 + n_local] = torch.from_numpy(st.pos)
        dist.all_reduce(mine)
        if rank == 0:
            full = orc.new_state(n_global)
            orc.reset(full)
            assert np.array_equal(mine.numpy(), full.pos)
        # (3) global advantage statistics: sum / sumsq / count all-reduce gives the unsharded mean/std
        g = torch.Generator().manual_seed(5)
        adv_all = torch.randn(64, generator=g, dtype=torch.float64)
        part = adv_all[rank::world_size]
        stats = torch.tensor([part.sum(), (part * part).sum(), float(part.numel())], dtype=torch.float64)
        dist.all_reduce(stats)
        mean = stats[0] / stats[2]
        var = (stats[1] - stats[2] * mean * mean) / (stats[2] - 1)
        assert torch.allclose(mean, adv_all.mean()) and torch.allclose(var.sqrt(), adv_all.std())
        Path(tmp, f"ok{rank}").write_text("ok")
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_sharding_and_allreduce(tmp_path):
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()

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


class _FakeEnv:
    possible_agents = ["cop_0", "cop_1", "thief_0"]
    num_envs, state_dim, device = 8, 1090, torch.device("cpu")

    class worlds:
        R = 90


def _kl_worker(rank: int, world_size: int, port: int, tmp: str):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world_size),
                      LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        from as_cops_and_thieves_b200 import mappo
        cfg = mappo.MAPPOConfig(rollouts=16, model="mlp", distributed=True, world_size=world_size, kl_threshold=0.01,
                                learning_epochs=3, mini_batches=4, random_timesteps=0, learning_starts=0,
                                policy_freeze_duration=0, opponent_freeze_duration=0)
        learner = mappo.MAPPOLearner(_FakeEnv(), cfg, seed=0)          # same initial weights on both ranks
        g = torch.Generator().manual_seed(100 + rank)                  # ... but DIFFERENT data
        T, N = cfg.rollouts, _FakeEnv.num_envs
        learner.mem_state.copy_(torch.rand((T, N, learner.n_value_in), generator=g))
        out = {}
        for a in learner.agents:
            m = learner.mem[a]
            m["obs"].copy_(torch.rand((T, N, learner.n_obs), generator=g))
            m["act"].copy_(torch.randint(0, 4, (T, N), generator=g))
            with torch.no_grad():
                logits, _ = learner.models[a]["policy"](m["obs"].view(T * N, 1, -1))
                logp = torch.log_softmax(logits.view(T, N, -1), -1).gather(-1, m["act"].unsqueeze(-1)).squeeze(-1)
            # rank 0 recorded exactly the current policy (KL = 0 at the first minibatch); rank 1's recorded log-probs are
            # far off, so ITS approximate KL is over the threshold from the very first minibatch of every epoch
            m["logp"].copy_(logp if rank == 0 else logp - 1.0)
            ret = torch.rand((T, N), generator=g)
            adv = torch.randn((T, N), generator=g)
            st = learner._ppo_update(a, ret, adv, mappo.UpdateStats())
            out[a] = st.minibatches
            # the decision was collective: no step was applied on EITHER rank, and the ranks stayed in lockstep
            assert st.minibatches == 0, (rank, a, st.minibatches)
            flat = torch.cat([p.detach().flatten() for p in learner.parameters(a)])
            other = flat.clone()
            dist.broadcast(other, src=0)
            assert torch.equal(flat, other), "weights diverged across ranks"
        # with the threshold off, both ranks apply every minibatch and still agree
        learner.cfg.kl_threshold = 0.0
        st = learner._ppo_update("cop_0", torch.rand((T, N), generator=g), torch.randn((T, N), generator=g), mappo.UpdateStats())
        assert st.minibatches == cfg.learning_epochs * cfg.mini_batches
        flat = torch.cat([p.detach().flatten() for p in learner.parameters("cop_0")])
        other = flat.clone()
        dist.broadcast(other, src=0)
        assert torch.equal(flat, other)
        Path(tmp, f"kl{rank}").write_text(str(out))
    finally:
        dist.destroy_process_group()


def test_two_rank_kl_early_stop_is_collective(tmp_path):
    """The round-1 learner broke out of the minibatch loop on the rank-LOCAL KL and then dead-locked (or silently
    paired mismatched buckets) in the gradient all-reduce.  Here rank 1 alone exceeds the threshold: both ranks must
    stop together, issue the same collectives, and keep identical weights."""
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_kl_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "kl0").exists() and (tmp_path / "kl1").exists()

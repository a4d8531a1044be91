#!/usr/bin/env python
"""MAPPO + PFSP self-play on the batched environment (BASELINE.json config 5): the batched counterpart of the
reference's ``self_play_driver.py``.  One process per GPU under torchrun; worlds are sharded, gradients
all-reduced over NCCL per minibatch.

    python tools/train_selfplay.py --map agh-map --worlds 4096 --iterations 2 --timesteps 256
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/train_selfplay.py ...

Prints one JSON line per self-play iteration (rank 0): environment agent-steps/s inside training, GAE / update /
all-reduce milliseconds per update, win rates.
"""
import argparse
import json
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from as_cops_and_thieves_b200 import selfplay  # noqa: E402
from as_cops_and_thieves_b200.env import BatchedCopsThievesEnv  # noqa: E402
from as_cops_and_thieves_b200.mappo import MAPPOConfig, MAPPOLearner  # noqa: E402
from as_cops_and_thieves_b200.maps import free_space_regions, load_named_map  # noqa: E402
from as_cops_and_thieves_b200.sharding import dist_env, shard_range  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--map", default="agh-map")
    ap.add_argument("--worlds", type=int, default=4096, help="worlds per GPU")
    ap.add_argument("--iterations", type=int, default=2)
    ap.add_argument("--timesteps", type=int, default=256, help="lockstep steps per self-play iteration")
    ap.add_argument("--rollouts", type=int, default=64)
    ap.add_argument("--model", default="lstm", choices=["lstm", "mlp"])
    ap.add_argument("--max-step-count", type=int, default=2000)     # self_play_driver.py:34
    ap.add_argument("--archive", default="policy_archive")
    ap.add_argument("--free-spawn", type=int, default=1)
    ap.add_argument("--autocast", default="none", choices=["none", "bf16"])
    ap.add_argument("--schedule", default="scaled", choices=["scaled", "reference", "off"],
                    help="random_timesteps / learning_starts / freeze durations: the reference's values divided by the "
                         "number of worlds (same number of transitions), the reference's values as they are, or none")
    a = ap.parse_args()

    rank, local_rank, world_size = dist_env()
    dev = torch.device(f"cuda:{local_rank}")
    torch.cuda.set_device(dev)
    if world_size > 1:
        import os
        import torch.distributed as dist
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    m = load_named_map(a.map)
    gid0, n_local = shard_range(a.worlds * world_size, rank, world_size)
    env = BatchedCopsThievesEnv(m, n_local, device=dev, seed=0, gid0=gid0, max_step_count=a.max_step_count,
                                spawn_override=free_space_regions(m) if a.free_spawn else None)
    # the reference's schedule (10 000 random / 15 000 before learning / 15 000 frozen policy) counts steps of ONE
    # environment; with thousands of lockstep worlds the same number of TRANSITIONS is reached in a few steps, so the
    # driver scales the durations by the world count unless told otherwise
    scale = max(1, a.worlds * world_size) if a.schedule == "scaled" else 1
    cfg = MAPPOConfig(rollouts=a.rollouts, model=a.model, distributed=world_size > 1, world_size=world_size,
                      update_autocast=a.autocast, random_timesteps=-(-10000 // scale) if a.schedule != "off" else 0,
                      learning_starts=-(-15000 // scale) if a.schedule != "off" else 0,
                      policy_freeze_duration=-(-15000 // scale) if a.schedule != "off" else 0,
                      opponent_freeze_duration=-(-15000 // scale) if a.schedule != "off" else 0)
    learner = MAPPOLearner(env, cfg, seed=0)
    archive = Path(a.archive) if rank == 0 else Path(a.archive + f".rank{rank}")
    for it in range(a.iterations):
        agg = {"gae_ms": 0.0, "update_ms": 0.0, "allreduce_ms": 0.0, "updates": 0}

        def cb(_l, stats):
            for st in stats.values():
                agg["gae_ms"] += st.gae_ms; agg["update_ms"] += st.update_ms; agg["allreduce_ms"] += st.allreduce_ms
            agg["updates"] += 1
        env_s0, t0 = learner.env_seconds, time.perf_counter()
        step0 = learner.timestep
        learner.train(learner.timestep + a.timesteps, callback=cb)
        train_s = time.perf_counter() - t0
        tc = selfplay.TrainingConfig
        res_c = selfplay.evaluate_agent(env, learner, tc.cop_role_prefix, tc.thief_role_prefix, archive / "thief", tc)
        res_t = selfplay.evaluate_agent(env, learner, tc.thief_role_prefix, tc.cop_role_prefix, archive / "cop", tc)
        ck = archive / f"joint_iter_{it}_full_agent.pt"
        archive.mkdir(parents=True, exist_ok=True)
        learner.save(str(ck))
        selfplay.add_policy_to_archive(str(ck), archive / "cop", it, "cop")
        selfplay.add_policy_to_archive(str(ck), archive / "thief", it, "thief")
        steps = learner.timestep - step0
        if rank == 0:
            u = max(agg["updates"], 1)
            print(json.dumps({
                "iteration": it, "n_gpus": world_size, "map": a.map, "worlds_per_gpu": a.worlds, "model": a.model,
                "parameters": learner.n_parameters(), "rollouts": a.rollouts, "lockstep_steps": steps,
                "rollout_agent_steps_per_s": a.worlds * world_size * len(learner.agents) * steps / max(learner.env_seconds - env_s0, 1e-9),
                "train_agent_steps_per_s": a.worlds * world_size * len(learner.agents) * steps / train_s,
                "gae_ms_per_update": agg["gae_ms"] / u, "ppo_ms_per_update": agg["update_ms"] / u,
                "allreduce_ms_per_update": agg["allreduce_ms"] / u,
                "eval_vs_archive": {"cop_trained": res_c, "thief_trained": res_t}}), flush=True)
    env.close()
    if world_size > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

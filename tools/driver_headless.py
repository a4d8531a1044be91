#!/usr/bin/env python
"""Headless counterpart of the reference's ``src/driver.py`` (``driver.py:53-91``): same objects, same calls —
``Map(mapfile)`` -> ``SimpleEnv(map, render_mode=...)`` -> ``env.reset()`` -> random ``Discrete(4)`` actions for
``env.agents`` -> ``env.step(actions)`` — minus the pygame window (``rgb_array`` frames are available through
``env.render()``).  ``--worlds N`` runs the same loop on N lockstep worlds through ``BatchedCopsThievesEnv``.

    python tools/driver_headless.py squarinth --steps 1000
    python tools/driver_headless.py /path/to/maps_templates/agh-map.json --worlds 4096 --steps 500
"""
import argparse
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from as_cops_and_thieves_b200 import BatchedCopsThievesEnv, Map, SimpleEnv, builtin_map_path  # noqa: E402


def main():
    ap = argparse.ArgumentParser(description="Cops and Robbers Game (headless, B200)")
    ap.add_argument("mapfile", help="a maps_templates JSON file, or the name of a built-in map")
    ap.add_argument("-r", "--render-mode", choices=["rgb_array"], default="rgb_array")
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--worlds", type=int, default=1)
    a = ap.parse_args()
    path = Path(a.mapfile)
    map = Map(str(path if path.exists() else builtin_map_path(a.mapfile)))
    if a.worlds == 1:
        env = SimpleEnv(map=map, map_image=None, render_mode=a.render_mode)
        observations, infos = env.reset()
        t0 = time.perf_counter()
        episodes = wins = 0
        for _ in range(a.steps):
            if not env.agents:                                   # episode over: driver.py would sit idle, a trainer resets
                observations, infos = env.reset()
            actions = {agent: env.action_space(agent).sample() for agent in env.agents}
            observations, rewards, terminations, _, infos = env.step(actions)
            if any(terminations.values()):
                episodes += 1
                wins += infos[next(iter(infos))]["winner"] == "cop"
        dt = time.perf_counter() - t0
        frame = env.render()
        print(f"{a.steps} steps, {episodes} episodes ({wins} won by the cops), last frame {getattr(frame, 'shape', None)}")
    else:
        env = BatchedCopsThievesEnv(map, a.worlds)
        env.reset()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        done = 0
        for _ in range(a.steps):
            actions = {agent: torch.randint(0, 4, (a.worlds, 1), device=env.device) for agent in env.agents}
            _, _, terminations, _, _ = env.step(actions)
            done += int(terminations[env.agents[0]].sum())
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(f"{a.steps} lockstep steps of {a.worlds} worlds, {done} episodes finished")
    print(f"{a.steps * a.worlds * len(env.possible_agents) / dt:.3e} agent-steps/s including Python")
    env.close()


if __name__ == "__main__":
    main()

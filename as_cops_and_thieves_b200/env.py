"""Reference-facing environment classes.

Two faces over the same CUDA step:

* ``BaseEnv`` / ``SimpleEnv`` — the reference's PettingZoo ``ParallelEnv`` surface for ONE world
  (``/root/reference/src/environments/base_env.py:30-554``, ``simple_env.py:6-58``): same
  constructor, ``reset`` / ``step`` / ``state`` / spaces / ``agents`` bookkeeping, numpy dict
  outputs with the reference's dtypes, so ``driver.py`` and ``evaluate_agents`` run against it.
* ``BatchedCopsThievesEnv`` — the surface skrl's multi-agent wrapper presents to
  ``SequentialTrainer`` / ``MAPPO`` (``wrap_env(env, wrapper="pettingzoo")`` at
  ``self_play_driver.py:35``), but with ``num_envs = N``: per-agent ``(N, 180)`` float32
  observations, ``(N, 1)`` rewards / flags, ``(N, 1090)`` state, auto-reset inside the step.

Both are thin: every number comes out of ``libcat_b200.so``.
"""
from __future__ import annotations

import functools
from typing import Dict, List, Mapping, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from . import spaces
from .maps import Map, compile_map
from .params import EnvParams, load_physical_params
from .worlds import CatWorlds

OBJECT_TYPE_MAX = 4  # ObjectType.EMPTY (utils/object_types.py:4-9)

# The reference's BaseEnv IS a pettingzoo.ParallelEnv (base_env.py:30), and skrl's wrap_env(env, wrapper="pettingzoo")
# (self_play_driver.py:35) checks for it.  Subclass the real class wherever pettingzoo is importable; this image has
# none, so the surface is duck-typed here and the base class is a plain object.
try:  # pragma: no cover - exercised only where pettingzoo exists
    from pettingzoo import ParallelEnv as _ParallelEnvBase  # type: ignore
    HAVE_PETTINGZOO = True
except Exception:
    _ParallelEnvBase = object
    HAVE_PETTINGZOO = False


def _agent_observation_space(n_rays: int, ray_length: float) -> "spaces.Dict":
    # entity.py:92-107
    return spaces.Dict({
        "distance": spaces.Box(low=0.0, high=ray_length, shape=(n_rays,), dtype=np.float16),
        "object_type": spaces.Box(low=0, high=OBJECT_TYPE_MAX, shape=(n_rays,), dtype=np.uint8),
    })


def _shared_observation_space(window: Tuple[float, float], ids: Sequence[str], n_cops: int, n_rays: int,
                              ray_length: float) -> "spaces.Dict":
    # observation_spaces.py:13-64
    max_dim = max(window)
    obs = _agent_observation_space(n_rays, ray_length)
    out = {}
    for team in (list(ids[:n_cops]), list(ids[n_cops:])):
        if not team:
            continue
        team_space = spaces.Dict({
            "own_obj_types": obs["object_type"], "own_distances": obs["distance"],
            "object_type_shared": obs["object_type"], "distance_shared": obs["distance"],
            "team_positions": spaces.Box(low=0.0, high=max_dim, shape=(len(team), 2), dtype=np.float16),
        })
        for aid in team:
            out[aid] = team_space
    return spaces.Dict(out)


class _EnvCommon:
    """Space/metadata plumbing shared by the single-world and the batched face."""

    metadata = {"render_modes": ["human", "rgb_array"]}  # base_env.py:49

    def _setup_common(self, map: Map, cmap, params: EnvParams) -> None:
        self.map = map
        self.width, self.height = map.window_dimensions
        self.possible_agents: List[str] = cmap.agent_ids                     # base_env.py:96
        self._agent_index = {a: i for i, a in enumerate(self.possible_agents)}
        obs_space = _agent_observation_space(params.n_rays, params.ray_length)
        self.observation_spaces = {a: obs_space for a in self.possible_agents}  # base_env.py:102-105
        self.action_spaces = {a: spaces.Discrete(4) for a in self.possible_agents}  # entity.py:88-90
        shared = _shared_observation_space(map.window_dimensions, self.possible_agents, cmap.n_cops,
                                           params.n_rays, params.ray_length)
        self.shared_observation_spaces = shared                              # base_env.py:113-115
        self._shared_observation_spaces = shared
        self.state_space = shared

    def observation_space(self, agent: str):
        return self.observation_spaces[agent]

    def action_space(self, agent: str):
        return self.action_spaces[agent]

    def get_base_observation_space_structure(self):  # base_env.py:243-253
        return self._shared_observation_spaces

    @functools.lru_cache(maxsize=None)
    def get_nested_agent_observation_spaces(self):  # base_env.py:256-284
        flat = {}
        for aid in self._shared_observation_spaces:
            sp = dict(self._shared_observation_spaces[aid].spaces.items())
            for other in self._shared_observation_spaces:
                if other != aid:
                    for key, s in self._shared_observation_spaces[other].spaces.items():
                        sp[f"{other}_{key}"] = s
            flat[aid] = spaces.Dict(sp)
        return spaces.Dict(flat)


def _make_params(max_step_count: int, time_step: float, physical: Optional[dict], **over) -> EnvParams:
    phys = dict(load_physical_params()) if physical is None else dict(physical)
    p = EnvParams(dt=float(time_step), max_step_count=int(max_step_count), **phys)
    for k, v in over.items():
        setattr(p, k, v)
    return p


class BaseEnv(_EnvCommon, _ParallelEnvBase):
    """Single-world PettingZoo ``ParallelEnv`` face (``base_env.py:30``): a real ``pettingzoo.ParallelEnv`` subclass
    when pettingzoo is importable (``HAVE_PETTINGZOO``), the same surface duck-typed otherwise.  Needs a CUDA device."""

    def __init__(self, map: Map, map_image=None, render_mode=None, max_step_count: int = 400,
                 time_step: float = 1 / 15.0, *, device: Union[str, torch.device] = "cuda:0",
                 physical_params: Optional[dict] = None, stale_shape_cache: bool = True):
        assert render_mode is None or render_mode in self.metadata["render_modes"]  # base_env.py:117
        self.render_mode = render_mode
        self.map_image = map_image
        self.max_step_count = max_step_count
        self.time_step = time_step
        self.step_count = 0
        params = _make_params(max_step_count, time_step, physical_params, auto_reset=0,
                              stale_shape_cache=int(stale_shape_cache))
        self._cmap = compile_map(map)
        self._params = params
        self._setup_common(map, self._cmap, params)
        # One world, consumed on the CPU every step: every output lives in mapped pinned host memory, so a step
        # is one kernel launch + one stream synchronisation, and the dicts below are built from numpy views.
        self._w = CatWorlds(self._cmap, 1, device=device, params=params, want_f32=False, want_shared=True,
                            pinned_outputs=True)
        self._acts = torch.zeros((1, len(self.possible_agents)), dtype=torch.uint8).pin_memory()
        self._acts_np = self._acts.numpy()
        w = self._w
        self._np = {k: getattr(w, k).numpy() for k in ("obs_dist", "obs_type", "reward", "terminated", "truncated",
                                                         "winner", "shared_dist", "shared_type", "team_pos")}
        self.agents: List[str] = []
        self._np_random_seed = None

    # ------------------------------------------------------------------ helpers
    def _observations(self) -> Dict[str, dict]:
        d, t = self._np["obs_dist"][0], self._np["obs_type"][0]
        return {a: {"distance": d[i].copy(), "object_type": t[i].copy()} for i, a in enumerate(self.possible_agents)}

    def _shared(self, obs: Dict[str, dict]) -> Dict[str, dict]:
        # observation_spaces.py:123-129 — team-mates share (alias) the merged arrays
        sd = self._np["shared_dist"][0].copy()
        st = self._np["shared_type"][0].copy()
        tp = self._np["team_pos"][0].copy()
        nc, na = self._cmap.n_cops, len(self.possible_agents)
        team_sd, team_st = [sd[0], sd[1]], [st[0], st[1]]      # one array object per team, shared by its members
        team_tp = [tp[0:nc], tp[nc:na]]
        out = {}
        for i, a in enumerate(self.possible_agents):
            team = 0 if i < nc else 1
            out[a] = {"own_obj_types": obs[a]["object_type"], "own_distances": obs[a]["distance"],
                      "object_type_shared": team_st[team], "distance_shared": team_sd[team],
                      "team_positions": team_tp[team]}
        return out

    # ------------------------------------------------------------------ PettingZoo API
    def reset(self, seed: Optional[int] = None, options: Optional[dict] = None):  # base_env.py:286-352
        if seed is not None:
            # base_env.py:307-311 re-creates np_random, so reset(seed=s) reproduces the same spawn every time:
            # re-key Philox AND restart the episode counter that is part of its key
            self._np_random_seed = seed
            self._w.set_seed(seed, restart_episodes=True)
        self.agents = self.possible_agents[:]
        self._w.reset()
        self._w.synchronize()
        observations = self._observations()
        infos = {a: {} for a in self.agents}
        self._state = self._shared(observations)
        self.step_count = 0
        return observations, infos

    def step(self, action: Mapping[str, int]):  # base_env.py:354-413
        self.step_count += 1
        if not action:
            self.agents = []
            return {}, {}, {}, {}, {}
        for i, a in enumerate(self.possible_agents):
            self._acts_np[0, i] = int(action[a])
        self._w.step(self._acts)
        self._w.synchronize()
        observations = self._observations()
        rew = self._np["reward"][0]
        terminated = bool(self._np["terminated"][0])
        truncated = bool(self._np["truncated"][0])
        winner_code = int(self._np["winner"][0])
        agents = self.agents
        rewards = {a: float(rew[self._agent_index[a]]) for a in agents}
        terminations = {a: terminated for a in agents}
        truncations = {a: truncated for a in agents}
        self._state = self._shared(observations)
        winner = None
        if terminated:
            self.agents = []
            winner = "cop" if winner_code == 0 else "thief"
        infos = {a: {"winner": winner} for a in agents}
        return {a: observations[a] for a in agents}, rewards, terminations, truncations, infos

    def state(self) -> dict:  # base_env.py:415-425
        return self._state

    def render(self):
        if self.render_mode == "rgb_array":
            from .render import render_rgb
            pos = self._w.get_state()["pos"][0].cpu().numpy()
            return render_rgb(self._cmap, pos)
        return None

    def close(self):
        self._w.close()

    def _get_info(self):  # base_env.py:459-475
        return {"step_count": self.step_count, "thief_count": self._cmap.n_thieves, "cop_count": self._cmap.n_cops}


class SimpleEnv(BaseEnv):
    """``simple_env.py:6-58``: the class every reference driver instantiates (dt 1/60, rgb_array)."""

    def __init__(self, map: Map, render_mode="rgb_array", map_image=None, max_step_count: int = 400,
                 time_step: float = 1 / 60.0, **kw):
        super().__init__(map=map, map_image=map_image, render_mode=render_mode, max_step_count=max_step_count,
                         time_step=time_step, **kw)

    def _get_info(self):
        info = super()._get_info()
        pos = self._w.get_state()["pos"][0].cpu().numpy()
        nc = self._cmap.n_cops
        info.update({"environment_type": "SimpleEnv",
                     "thief_positions": [tuple(int(v) for v in p) for p in pos[nc:]],
                     "cop_positions": [tuple(int(v) for v in p) for p in pos[:nc]]})
        return info


class BatchedCopsThievesEnv(_EnvCommon, _ParallelEnvBase):
    """N worlds behind the surface skrl's multi-agent trainers use.

    ``agents`` never empties: finished worlds are re-spawned inside the step kernel and report the
    observation of the new episode, with the terminal reward / flags of the old one.
    """

    def __init__(self, map: Map, num_envs: int, *, max_step_count: int = 400, time_step: float = 1 / 60.0,
                 device: Union[str, torch.device] = "cuda:0", seed: int = 0, gid0: int = 0,
                 physical_params: Optional[dict] = None, stale_shape_cache: bool = True,
                 spawn_override: Optional[dict] = None, cell: Optional[float] = None, render_mode=None):
        params = _make_params(max_step_count, time_step, physical_params, auto_reset=1, seed=int(seed),
                              stale_shape_cache=int(stale_shape_cache))
        self._cmap = compile_map(map, spawn_override=spawn_override, cell=cell)
        self._params = params
        self._setup_common(map, self._cmap, params)
        self._w = CatWorlds(self._cmap, num_envs, device=device, gid0=gid0, params=params, want_f32=True, want_critic=True)
        self.num_envs = int(num_envs)
        self.num_agents = len(self.possible_agents)
        self.agents = self.possible_agents[:]
        self.device = self._w.device
        self.render_mode = render_mode
        self.max_step_count = max_step_count
        self.state_spaces = {a: self.shared_observation_spaces for a in self.possible_agents}
        self.state_dim = self._w.S

    @property
    def worlds(self) -> CatWorlds:
        return self._w

    def state_space_of(self, agent: str):
        return self.state_spaces[agent]

    def _obs_dict(self) -> Dict[str, torch.Tensor]:
        return self._views()[0]

    def _views(self):
        """The step's return values are fixed views of the kernel's output buffers, built once: per-agent
        observation / reward slices, and the u8 flags reinterpreted as bool (the kernel writes 0 / 1), so a step
        costs one kernel launch and no torch ops.  The same dict objects are returned by every call; their
        tensors are overwritten by the next step (copy what must outlive it, as skrl's memories do)."""
        v = getattr(self, "_cached_views", None)
        if v is None:
            w = self._w
            obs = {a: w.obs_f32[i] for i, a in enumerate(self.possible_agents)}
            rew = {a: w.reward[:, i:i + 1] for i, a in enumerate(self.possible_agents)}
            term = w.terminated.view(torch.bool).unsqueeze(1)
            trunc = w.truncated.view(torch.bool).unsqueeze(1)
            v = (obs, rew, {a: term for a in self.possible_agents}, {a: trunc for a in self.possible_agents},
                 {a: {"winner": w.winner} for a in self.possible_agents})
            self._cached_views = v
        return v

    def reset(self, seed: Optional[int] = None, options: Optional[dict] = None):
        if seed is not None:
            self._w.set_seed(seed, restart_episodes=True)
        self._w.reset()
        return self._obs_dict(), {a: {} for a in self.possible_agents}

    def step(self, actions: Union[Mapping[str, torch.Tensor], torch.Tensor]):
        if isinstance(actions, Mapping):
            acts = [actions[a].reshape(-1) for a in self.possible_agents]
            if all(t.dtype == torch.int64 and t.is_contiguous() and t.device == self.device for t in acts):
                self._w.step(acts)
            else:
                self._w.step(torch.stack([t.to(self.device) for t in acts], dim=1).to(torch.uint8).contiguous())
        else:
            self._w.step(actions)
        return self._views()

    def state(self) -> torch.Tensor:
        return self._w.state_f32

    def critic(self) -> torch.Tensor:
        """The critic's front end, written by the step kernel itself: the 4 ray channels ``LSTMValue`` cuts out of
        ``state()`` (``lstm_value_net.py:122-137``) as ``(N, 4, R)`` float32 — no slice / stack copies on the way."""
        return self._w.critic_f32

    def render(self, world: int = 0):
        from .render import render_rgb
        pos = self._w.get_state()["pos"][world].cpu().numpy()
        return render_rgb(self._cmap, pos)

    def close(self) -> None:
        self._w.close()

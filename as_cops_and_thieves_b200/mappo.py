"""Minimal MAPPO learner over the batched environment (BASELINE config 5).

skrl is not installable in this image, so this restates the part of ``skrl.multi_agents.torch.mappo.MAPPO``
the reference drives (``/root/reference/src/training/orchestration.py:133-142``,
``/root/reference/src/utils/agent_learning_utils.py:172-199``) with the reference's hyper-parameters
(``/root/reference/src/configs/mappo_config.py:5-50``) and model architectures
(``/root/reference/src/models/policy_net.py:17-33``, ``value_net.py:18-28``) in plain PyTorch:

* rollout memory on the device, ``rollouts`` steps x N worlds, one policy + one centralised critic per agent;
* GAE + advantage normalisation through the CUDA kernels (``gae.compute_gae``);
* PPO-clip surrogate, entropy bonus, scaled value loss, KL early stop, joint grad-norm clip;
* under ``torchrun`` the gradients of every minibatch are all-reduced over NCCL in one flat bucket
  (``sharding.allreduce_gradients``) — the only collective in the system besides the optional advantage
  statistics.

The dense nets are library PyTorch on purpose (out of scope for kernels).
"""
from __future__ import annotations

import itertools
import time
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import torch
import torch.nn as nn

from .gae import compute_gae
from .sharding import allreduce_gradients


class PolicyNet(nn.Module):
    """``Policy`` of the reference (``policy_net.py:17-33``): Conv1d(2→64,k5,s2) → Conv1d(64→32,k5,s3) →
    Linear(416→256) → MLP → 4 logits, on the observation viewed as (B, 2, 90) = [distance | object_type]."""

    def __init__(self, n_obs: int = 180, n_actions: int = 4):
        super().__init__()
        self.len_ch = n_obs // 2
        l1 = (self.len_ch - 5) // 2 + 1
        l2 = (l1 - 5) // 3 + 1
        self.features_extractor = nn.Sequential(
            nn.Conv1d(2, 64, kernel_size=5, stride=2), nn.ReLU(),
            nn.Conv1d(64, 32, kernel_size=5, stride=3), nn.ReLU(),
            nn.Flatten(), nn.Linear(32 * l2, 256), nn.Tanh())
        self.net = nn.Sequential(nn.Linear(256, 128), nn.ReLU(), nn.Linear(128, 64), nn.ReLU(), nn.Linear(64, n_actions))

    def forward(self, obs: torch.Tensor) -> torch.Tensor:
        return self.net(self.features_extractor(obs.view(obs.size(0), 2, self.len_ch)))


class ValueNet(nn.Module):
    """``Value`` of the reference (``value_net.py:18-28``): MLP 1090→512→256→128→64→1 on the flattened state."""

    def __init__(self, n_state: int = 1090):
        super().__init__()
        self.net = nn.Sequential(nn.Linear(n_state, 512), nn.ReLU(), nn.Linear(512, 256), nn.ReLU(),
                                 nn.Linear(256, 128), nn.ReLU(), nn.Linear(128, 64), nn.ReLU(), nn.Linear(64, 1))

    def forward(self, state: torch.Tensor) -> torch.Tensor:
        return self.net(state)


@dataclass
class MAPPOConfig:
    rollouts: int = 64                 # steps per update (reference: 4096 with ONE env; here x N worlds)
    learning_epochs: int = 4           # mappo_config.py:44
    mini_batches: int = 4              # :45
    discount_factor: float = 0.99      # skrl default
    lambda_: float = 0.95              # skrl default
    learning_rate: float = 1e-4        # :47
    ratio_clip: float = 0.15           # :48
    entropy_loss_scale: float = 0.02   # :46
    value_loss_scale: float = 0.5      # :13
    grad_norm_clip: float = 0.5        # :14
    kl_threshold: float = 0.015        # :11
    distributed: bool = False
    world_size: int = 1


@dataclass
class UpdateStats:
    policy_loss: float = 0.0
    value_loss: float = 0.0
    entropy: float = 0.0
    kl: float = 0.0
    minibatches: int = 0
    gae_ms: float = 0.0
    allreduce_ms: float = 0.0
    update_ms: float = 0.0


def ppo_losses(logits: torch.Tensor, actions: torch.Tensor, old_log_prob: torch.Tensor, advantages: torch.Tensor,
               values: torch.Tensor, returns: torch.Tensor, cfg: MAPPOConfig):
    """The three loss terms of skrl's ``MAPPO._update`` for one minibatch, plus the approximate KL it uses
    for early stopping.  Pure function of tensors (unit-tested on CPU)."""
    dist = torch.distributions.Categorical(logits=logits)
    new_log_prob = dist.log_prob(actions)
    log_ratio = new_log_prob - old_log_prob
    kl = ((torch.exp(log_ratio) - 1.0) - log_ratio).mean()
    ratio = torch.exp(log_ratio)
    surrogate = advantages * ratio
    clipped = advantages * torch.clip(ratio, 1.0 - cfg.ratio_clip, 1.0 + cfg.ratio_clip)
    policy_loss = -torch.min(surrogate, clipped).mean()
    entropy = dist.entropy().mean()
    entropy_loss = -cfg.entropy_loss_scale * entropy
    value_loss = cfg.value_loss_scale * torch.nn.functional.mse_loss(values, returns)
    return policy_loss, entropy_loss, value_loss, kl, entropy


class MAPPOLearner:
    def __init__(self, env, cfg: Optional[MAPPOConfig] = None, seed: int = 0):
        self.env = env
        self.cfg = cfg or MAPPOConfig()
        self.device = env.device
        self.agents: List[str] = list(env.possible_agents)
        n_obs = 2 * env.worlds.R
        torch.manual_seed(seed)     # identical initial weights on every rank
        self.policies = {a: PolicyNet(n_obs, 4).to(self.device) for a in self.agents}
        self.values = {a: ValueNet(env.state_dim).to(self.device) for a in self.agents}
        self.optimizers = {a: torch.optim.Adam(itertools.chain(self.policies[a].parameters(),
                                                               self.values[a].parameters()), lr=self.cfg.learning_rate)
                           for a in self.agents}
        T, N = self.cfg.rollouts, env.num_envs
        dev = self.device
        self.mem = {a: dict(obs=torch.zeros((T, N, n_obs), device=dev), act=torch.zeros((T, N), dtype=torch.int64, device=dev),
                            logp=torch.zeros((T, N), device=dev), rew=torch.zeros((T, N), device=dev),
                            val=torch.zeros((T, N), device=dev)) for a in self.agents}
        self.mem_state = torch.zeros((T, N, env.state_dim), device=dev)
        self.mem_done = torch.zeros((T, N), dtype=torch.bool, device=dev)
        self._obs = None
        self.env_steps = 0
        self.env_seconds = 0.0

    def n_parameters(self) -> int:
        return sum(p.numel() for a in self.agents for p in itertools.chain(self.policies[a].parameters(),
                                                                         self.values[a].parameters()))

    @torch.no_grad()
    def collect(self) -> None:
        """``rollouts`` lockstep transitions of all worlds with the current policies (sampled actions)."""
        env, cfg = self.env, self.cfg
        if self._obs is None:
            self._obs, _ = env.reset()
        t0 = time.perf_counter()
        for t in range(cfg.rollouts):
            state = env.state()
            self.mem_state[t].copy_(state)
            actions = {}
            for a in self.agents:
                o = self._obs[a]
                self.mem[a]["obs"][t].copy_(o)
                dist = torch.distributions.Categorical(logits=self.policies[a](o))
                act = dist.sample()
                self.mem[a]["act"][t] = act
                self.mem[a]["logp"][t] = dist.log_prob(act)
                self.mem[a]["val"][t] = self.values[a](state).squeeze(-1)
                actions[a] = act
            obs, rew, term, trunc, _ = env.step(actions)
            for a in self.agents:
                self.mem[a]["rew"][t] = rew[a].squeeze(-1)
            self.mem_done[t] = (term[self.agents[0]] | trunc[self.agents[0]]).squeeze(-1)
            self._obs = obs          # views of the env's output buffers: consumed before the next step
        torch.cuda.synchronize(self.device)
        self.env_seconds += time.perf_counter() - t0
        self.env_steps += cfg.rollouts

    def update(self) -> Dict[str, UpdateStats]:
        cfg = self.cfg
        stats = {}
        T, N = cfg.rollouts, self.env.num_envs
        last_state = self.env.state()
        for a in self.agents:
            st = UpdateStats()
            ev0, ev1, ev2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            ev0.record()
            mem = self.mem[a]
            with torch.no_grad():
                last_values = self.values[a](last_state).squeeze(-1)
            returns, advantages = compute_gae(mem["rew"], self.mem_done, mem["val"], last_values, cfg.discount_factor,
                                              cfg.lambda_, normalize=True, distributed=cfg.distributed)
            ev1.record()
            obs = mem["obs"].view(T * N, -1)
            state = self.mem_state.view(T * N, -1)
            act, logp = mem["act"].view(-1), mem["logp"].view(-1)
            ret, adv = returns.view(-1), advantages.view(-1)
            params = list(itertools.chain(self.policies[a].parameters(), self.values[a].parameters()))
            ar_ms = 0.0
            stop = False
            for _epoch in range(cfg.learning_epochs):
                perm = torch.randperm(T * N, device=self.device)
                for idx in perm.chunk(cfg.mini_batches):
                    logits = self.policies[a](obs[idx])
                    values = self.values[a](state[idx]).squeeze(-1)
                    pl, el, vl, kl, ent = ppo_losses(logits, act[idx], logp[idx], adv[idx], values, ret[idx], cfg)
                    if cfg.kl_threshold and float(kl) > cfg.kl_threshold:
                        stop = True
                        break
                    self.optimizers[a].zero_grad(set_to_none=False)
                    (pl + el + vl).backward()
                    if cfg.distributed:
                        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        e0.record()
                        allreduce_gradients(params, cfg.world_size)
                        e1.record()
                        e1.synchronize()
                        ar_ms += e0.elapsed_time(e1)
                    if cfg.grad_norm_clip > 0:
                        nn.utils.clip_grad_norm_(params, cfg.grad_norm_clip)
                    self.optimizers[a].step()
                    st.policy_loss, st.value_loss, st.entropy, st.kl = float(pl), float(vl), float(ent), float(kl)
                    st.minibatches += 1
                if stop:
                    break
            ev2.record()
            ev2.synchronize()
            st.gae_ms, st.update_ms, st.allreduce_ms = ev0.elapsed_time(ev1), ev1.elapsed_time(ev2), ar_ms
            stats[a] = st
        return stats

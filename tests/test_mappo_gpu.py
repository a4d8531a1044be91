"""MAPPO learner + PFSP self-play iteration on the batched CUDA environment."""
import json
from pathlib import Path
from types import SimpleNamespace

import pytest
import torch

from as_cops_and_thieves_b200.env import BatchedCopsThievesEnv
from as_cops_and_thieves_b200.maps import load_named_map
from as_cops_and_thieves_b200.mappo import MAPPOConfig, MAPPOLearner
from as_cops_and_thieves_b200 import selfplay

pytestmark = pytest.mark.gpu

#: learn from the first rollout (the reference's defaults wait 15 000 timesteps: test_reference_schedule below)
NOW = dict(random_timesteps=0, learning_starts=0, policy_freeze_duration=0, opponent_freeze_duration=0)


@pytest.mark.parametrize("kind,autocast", [("lstm", "none"), ("mlp", "none"), ("lstm", "bf16")])
def test_learner_collects_and_updates(cuda_device, kind, autocast):
    env = BatchedCopsThievesEnv(load_named_map("squarinth"), 128, device=cuda_device, seed=1, max_step_count=40)
    cfg = MAPPOConfig(rollouts=32, model=kind, kl_threshold=0.0, update_autocast=autocast, **NOW)
    learner = MAPPOLearner(env, cfg, seed=0)
    assert learner.n_parameters() == (2_588_175 if kind == "lstm" else learner.n_parameters())
    before = {a: [p.detach().clone() for p in learner.parameters(a)] for a in learner.agents}
    for _ in range(2):
        learner.collect()
        stats = learner.update()
    assert learner.timestep == 64
    assert set(stats) == set(env.possible_agents)
    for a, st in stats.items():
        assert st.minibatches == cfg.learning_epochs * cfg.mini_batches
        for v in (st.policy_loss, st.value_loss, st.entropy, st.kl):
            assert v == v and abs(v) < 1e6      # finite
        assert 0.0 < st.entropy <= 1.3863 + 1e-4  # <= ln 4
        assert any(not torch.equal(p0, p1) for p0, p1 in zip(before[a], learner.parameters(a)))
    # episodes ended inside the rollout (max_step_count=40 < 64): resets were recorded and used
    assert bool(learner.mem_done.any()) and bool(learner.mem_reset.any())
    # frozen policies (train_simultaneously_and_evaluate freezes them first): only the critic moves
    learner.freeze("cop_0", "policy", True)
    pol_before = [p.detach().clone() for p in learner.models["cop_0"]["policy"].parameters()]
    val_before = [p.detach().clone() for p in learner.models["cop_0"]["value"].parameters()]
    learner.collect()
    learner.update()
    assert all(torch.equal(a, b) for a, b in zip(pol_before, learner.models["cop_0"]["policy"].parameters()))
    assert any(not torch.equal(a, b) for a, b in zip(val_before, learner.models["cop_0"]["value"].parameters()))
    env.close()


def test_batched_evaluation_and_self_play_iteration(cuda_device, tmp_path):
    env = BatchedCopsThievesEnv(load_named_map("squarinth"), 256, device=cuda_device, seed=2, max_step_count=30)
    learner = MAPPOLearner(env, MAPPOConfig(rollouts=16, model="lstm", **NOW), seed=0)
    cop, thief = selfplay.evaluate_agents(env, learner, n_episodes=2)
    assert 0.0 <= cop <= 1.0 and 0.0 <= thief <= 1.0 and cop + thief == pytest.approx(1.0)   # every episode has a winner
    tc = SimpleNamespace(**{**vars(selfplay.TrainingConfig), "n_trial_episodes": 1, "num_self_play_iterations": 2})
    ck0 = selfplay.self_play_iteration(env, learner, 0, tmp_path, timesteps=16, training_config=tc)
    assert Path(ck0).name == "joint_iter_0_full_agent.pt"
    assert (tmp_path / "cop" / "cop_iter_0.pt").exists() and (tmp_path / "thief" / "thief_iter_0.pt").exists()
    ck1 = selfplay.self_play_iteration(env, learner, 1, tmp_path, timesteps=16, training_config=tc)
    assert learner.timestep == 32
    # the second iteration evaluated both roles against iteration 0's archive and recorded the outcomes
    for role in ("cop", "thief"):
        table = json.loads((tmp_path / role / "win_rates.json").read_text())
        assert table[f"{role}_iter_0.pt"]["games"] == 1
    # a checkpoint restores the exact weights, and role-wise loading leaves the other role alone
    other = MAPPOLearner(env, MAPPOConfig(rollouts=16, model="lstm", **NOW), seed=5)
    thief_before = [p.detach().clone() for p in other.parameters("thief_0")]
    other.load(ck1, role_prefix="cop")
    assert all(torch.equal(a, b) for a, b in zip(learner.parameters("cop_1"), other.parameters("cop_1")))
    assert all(torch.equal(a, b) for a, b in zip(thief_before, other.parameters("thief_0")))
    env.close()


def test_reference_schedule_and_device_side_kl_stop(cuda_device):
    """CFG_AGENT / CFG_TRAINER semantics (mappo_config.py:5-63, agent_learning_utils.py:188-199, README.md:78-152):
    uniform random actions for `random_timesteps`, no update before `learning_starts`, every policy frozen until
    `policy_freeze_duration`; and skrl's KL early stop taken on the device (no minibatch applied once the KL exceeds
    the threshold, for the rest of that epoch)."""
    d = MAPPOConfig()
    assert (d.random_timesteps, d.learning_starts, d.policy_freeze_duration, d.opponent_freeze_duration) == (10000, 15000, 15000, 15000)
    env = BatchedCopsThievesEnv(load_named_map("squarinth"), 64, device=cuda_device, seed=3, max_step_count=40)
    cfg = MAPPOConfig(rollouts=16, model="mlp", random_timesteps=16, learning_starts=32, policy_freeze_duration=48,
                      opponent_freeze_duration=48, kl_threshold=0.0)
    learner = MAPPOLearner(env, cfg, seed=0)
    log = []
    pol0 = [p.detach().clone() for p in learner.models["cop_0"]["policy"].parameters()]

    def cb(l, stats):
        log.append((l.timestep, dict(l.frozen["cop_0"]), sorted(stats), {a: s.minibatches for a, s in stats.items()}))
    learner.train(80, callback=cb)
    # rollout 1 (t=16): below learning_starts -> no update; rollout 2 (t=32): updates start, policies frozen (value only);
    # rollout 3 (t=48): policy_freeze_duration reached -> policies unfrozen before that update
    assert log[0][0] == 16 and log[0][2] == []
    assert log[1][0] == 32 and log[1][1] == {"policy": True, "value": False} and log[1][2] == sorted(env.possible_agents)
    assert log[2][0] == 48 and log[2][1] == {"policy": False, "value": False}
    assert all(torch.equal(a, b) for a, b in zip(pol0, learner.models["cop_0"]["policy"].parameters())) is False
    # the frozen phase really left the policy alone: replay it
    l2 = MAPPOLearner(env, cfg, seed=0)
    pol0 = [p.detach().clone() for p in l2.models["cop_0"]["policy"].parameters()]
    l2.train(32)
    assert all(torch.equal(a, b) for a, b in zip(pol0, l2.models["cop_0"]["policy"].parameters()))
    # KL stop: an absurdly small threshold stops every epoch after its first minibatch (the first one of an update has
    # KL = 0 exactly — same weights as the rollout — so exactly one step is applied, in the first epoch)
    cfg3 = MAPPOConfig(rollouts=16, model="mlp", kl_threshold=1e-12, **NOW)
    l3 = MAPPOLearner(env, cfg3, seed=0)
    l3.collect()
    st3 = l3.update()
    for a, s in st3.items():
        assert 1 <= s.minibatches < cfg3.learning_epochs * cfg3.mini_batches, (a, s.minibatches)
    env.close()

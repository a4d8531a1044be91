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


@pytest.mark.parametrize("kind,autocast", [("lstm", "none"), ("mlp", "none"), ("lstm", "bf16")])
def test_learner_collects_and_updates(cuda_device, kind, autocast):
    env = BatchedCopsThievesEnv(load_named_map("squarinth"), 128, device=cuda_device, seed=1, max_step_count=40)
    cfg = MAPPOConfig(rollouts=32, model=kind, kl_threshold=0.0, update_autocast=autocast)
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
    learner = MAPPOLearner(env, MAPPOConfig(rollouts=16, model="lstm"), seed=0)
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
    other = MAPPOLearner(env, MAPPOConfig(rollouts=16, model="lstm"), seed=5)
    thief_before = [p.detach().clone() for p in other.parameters("thief_0")]
    other.load(ck1, role_prefix="cop")
    assert all(torch.equal(a, b) for a, b in zip(learner.parameters("cop_1"), other.parameters("cop_1")))
    assert all(torch.equal(a, b) for a, b in zip(thief_before, other.parameters("thief_0")))
    env.close()

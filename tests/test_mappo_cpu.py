"""Host logic of the MAPPO learner and the PFSP self-play helpers — CPU only."""
import json
import random
from collections import deque
from pathlib import Path

import pytest
import torch

from as_cops_and_thieves_b200 import mappo, selfplay


def test_model_sizes_match_the_reference_architectures():
    # LSTMPolicy 340,388 + LSTMValue 522,337 per agent, 2,588,175 for three agents
    pol, val = mappo.LSTMPolicyNet(180, 4), mappo.LSTMValueNet(1090)
    assert mappo.n_parameters(pol) == 340_388
    assert mappo.n_parameters(val) == 522_337
    models = mappo.build_models(["cop_0", "cop_1", "thief_0"], 180, 1090, "lstm", "cpu")
    assert sum(mappo.n_parameters(m) for a in models.values() for m in a.values()) == 2_588_175
    mlp = mappo.build_models(["cop_0"], 180, 1090, "mlp", "cpu")["cop_0"]
    logits, _ = mlp["policy"](torch.zeros(3, 2, 180))
    value, _ = mlp["value"](torch.zeros(3, 2, 1090))
    assert logits.shape == (3, 2, 4) and value.shape == (3, 2)
    with pytest.raises(ValueError):
        mappo.build_models(["cop_0"], 180, 1090, "transformer", "cpu")


def test_critic_channels_follow_the_reference_slices():
    # lstm_value_net.py:124-137: channels = [own_obj_types 270:360, own_distances 180:270, type_shared 90:180, dist_shared 0:90]
    net = mappo.LSTMValueNet(1090)
    state = torch.arange(1090, dtype=torch.float32).repeat(2, 1)
    ch = net.critic_channels(state)
    assert ch.shape == (2, 4, 90)
    assert ch[0, 0, 0] == 270 and ch[0, 1, 0] == 180 and ch[0, 2, 0] == 90 and ch[0, 3, 0] == 0 and ch[0, 3, 89] == 89


@pytest.mark.parametrize("net_cls,width", [(mappo.LSTMPolicyNet, 180), (mappo.LSTMValueNet, 1090)])
def test_sequence_with_resets_equals_step_by_step(net_cls, width):
    """One call over a 16-step sequence with mid-sequence episode ends must equal feeding the steps one at a
    time and zeroing the state of finished rows — the rollout does the latter, the PPO update the former."""
    torch.manual_seed(0)
    net = net_cls(width).double()
    B, L = 5, 16
    x = torch.rand(B, L, width, dtype=torch.float64)
    reset = torch.zeros(B, L, dtype=torch.bool)
    reset[0, 0] = reset[1, 3] = reset[1, 9] = reset[4, 15] = reset[2, 9] = True
    h0 = tuple(torch.randn(net.lstm.num_layers, B, net.lstm.hidden_size, dtype=torch.float64) for _ in range(2))
    full, (hf, cf) = net(x, h0, reset)
    hc = h0
    steps = []
    for t in range(L):
        o, hc = net(x[:, t:t + 1], hc, reset[:, t:t + 1])
        steps.append(o)
    torch.testing.assert_close(full, torch.cat(steps, dim=1), rtol=1e-10, atol=1e-12)
    torch.testing.assert_close(hf, hc[0], rtol=1e-10, atol=1e-12)
    none, _ = net(x, h0, None)
    same, _ = net(x, h0, torch.zeros(B, L, dtype=torch.bool))
    torch.testing.assert_close(none, same)
    assert not torch.allclose(none[1], full[1])


def test_ppo_losses_match_the_skrl_formulas():
    torch.manual_seed(1)
    cfg = mappo.MAPPOConfig()
    n = 64
    logits = torch.randn(n, 4, dtype=torch.float64)
    actions = torch.randint(0, 4, (n,))
    old = torch.log_softmax(torch.randn(n, 4, dtype=torch.float64), -1).gather(1, actions[:, None]).squeeze(1)
    adv, values, returns = (torch.randn(n, dtype=torch.float64) for _ in range(3))
    pl, el, vl, kl, ent = mappo.ppo_losses(logits, actions, old, adv, values, returns, cfg)
    logp = torch.log_softmax(logits, -1)
    new = logp.gather(1, actions[:, None]).squeeze(1)
    ratio = (new - old).exp()
    want_pl = -torch.minimum(adv * ratio, adv * ratio.clamp(1 - 0.15, 1 + 0.15)).mean()
    want_ent = -(logp.exp() * logp).sum(-1).mean()
    torch.testing.assert_close(pl, want_pl)
    torch.testing.assert_close(el, -0.02 * want_ent)
    torch.testing.assert_close(vl, 0.5 * ((values - returns) ** 2).mean())
    torch.testing.assert_close(kl, ((ratio - 1) - (new - old)).mean())
    # the reference's agent configuration (mappo_config.py:5-50)
    assert (cfg.learning_epochs, cfg.mini_batches, cfg.learning_rate, cfg.ratio_clip) == (4, 4, 1e-4, 0.15)
    assert (cfg.kl_threshold, cfg.value_loss_scale, cfg.grad_norm_clip, cfg.entropy_loss_scale) == (0.015, 0.5, 0.5, 0.02)


def test_win_rate_table_window_and_json_format(tmp_path):
    arch = tmp_path / "thief"
    for outcome in [True, True, False, True]:
        selfplay.update_policy_win_rate(arch, "thief_iter_0.pt", outcome, buffer_size=3)
    raw = json.loads((arch / "win_rates.json").read_text())["thief_iter_0.pt"]
    assert raw == {"wins": 3, "games": 4, "recent_outcomes": [1, 0, 1], "buffer_size": 3}
    table = selfplay.load_win_rates(arch)
    assert isinstance(table["thief_iter_0.pt"]["recent_outcomes"], deque)
    assert selfplay.current_win_rate(table["thief_iter_0.pt"]) == pytest.approx(2 / 3)
    assert selfplay.current_win_rate(None) == 0.5
    assert selfplay.current_win_rate({"wins": 1, "games": 4, "recent_outcomes": []}) == 0.25
    (arch / "win_rates.json").write_text("{not json")
    assert selfplay.load_win_rates(arch) == {}
    assert selfplay.load_win_rates(tmp_path / "missing") == {}


def test_pfsp_weight_and_sampling(tmp_path):
    assert selfplay.pfsp_weight(0.5) == 1.0
    assert selfplay.pfsp_weight(0.75) == pytest.approx(0.5)
    assert selfplay.pfsp_weight(1.0) == 1e-3 and selfplay.pfsp_weight(0.0) == 1e-3
    arch = tmp_path / "cop"
    ck = tmp_path / "joint.pt"
    ck.write_bytes(b"x")
    assert selfplay.sample_policy_from_archive(arch, "cop", "pfsp") is None
    assert selfplay.get_latest_policy_from_archive(arch, "cop") is None
    for it in (0, 2, 10):
        dst = selfplay.add_policy_to_archive(str(ck), arch, it, "cop")
        assert dst.name == f"cop_iter_{it}.pt"
    assert Path(selfplay.get_latest_policy_from_archive(arch, "cop")).name == "cop_iter_10.pt"   # numeric, not lexicographic
    assert Path(selfplay.sample_policy_from_archive(arch, "cop", "latest")).name == "cop_iter_10.pt"
    assert Path(selfplay.sample_policy_from_archive(arch, "cop", "no-such-strategy")).name == "cop_iter_10.pt"
    for _ in range(20):                       # cop_iter_0 always wins -> weight 1e-3; cop_iter_2 at 50 % -> weight 1
        selfplay.update_policy_win_rate(arch, "cop_iter_0.pt", True, 20)
    for i in range(20):
        selfplay.update_policy_win_rate(arch, "cop_iter_2.pt", i % 2 == 0, 20)
    rng = random.Random(0)
    picks = [Path(selfplay.sample_policy_from_archive(arch, "cop", "pfsp", rng)).name for _ in range(2000)]
    frac = {n: picks.count(n) / len(picks) for n in ("cop_iter_0.pt", "cop_iter_2.pt", "cop_iter_10.pt")}
    assert frac["cop_iter_0.pt"] < 0.01                      # weights 0.001 : 1 : 1 (unseen policy defaults to 0.5)
    assert abs(frac["cop_iter_2.pt"] - 0.5) < 0.05 and abs(frac["cop_iter_10.pt"] - 0.5) < 0.05
    assert {Path(selfplay.sample_policy_from_archive(arch, "cop", "random", rng)).name for _ in range(200)} == set(frac)


class _FakeEnv:
    """Just enough of BatchedCopsThievesEnv to construct a learner on the CPU (no rollouts, no GAE kernels)."""
    possible_agents = ["cop_0", "cop_1", "thief_0"]
    num_envs, state_dim, device = 4, 1090, torch.device("cpu")

    class worlds:
        R = 90


def test_flat_gradient_bucket_and_freezing():
    learner = mappo.MAPPOLearner(_FakeEnv(), mappo.MAPPOConfig(rollouts=16, model="mlp"), seed=0)
    a = "cop_0"
    flat = learner._flat_grad[a]
    params = learner.parameters(a)
    assert flat.numel() == sum(p.numel() for p in params)
    off = 0
    for p in params:                                   # every .grad is a view into the agent's flat bucket, in order
        assert p.grad.data_ptr() == flat.data_ptr() + 4 * off and p.grad.shape == p.shape
        off += p.numel()
    logits, _ = learner.models[a]["policy"](torch.rand(4, 2, 180))
    value, _ = learner.models[a]["value"](torch.rand(4, 2, 1090))
    (logits.sum() + value.sum()).backward()
    assert float(flat.abs().sum()) > 0                 # autograd accumulated straight into the bucket
    for p in params:
        assert p.grad.data_ptr() >= flat.data_ptr() and p.grad.data_ptr() < flat.data_ptr() + 4 * flat.numel()
    learner.optimizers[a].zero_grad(set_to_none=False)
    assert float(flat.abs().sum()) == 0
    # frozen networks get no gradient at all (Adam then skips them), trainable ones keep their slice
    learner.freeze(a, "policy", True)
    assert all(p.grad is None and not p.requires_grad for p in learner.models[a]["policy"].parameters())
    assert all(p.grad is not None for p in learner.models[a]["value"].parameters())
    learner.freeze(a, "policy", False)
    off = 0
    for p in params:
        assert p.requires_grad and p.grad.data_ptr() == flat.data_ptr() + 4 * off
        off += p.numel()
    # identical initial weights for a given seed (every rank builds the same nets before the first all-reduce)
    other = mappo.MAPPOLearner(_FakeEnv(), mappo.MAPPOConfig(rollouts=16, model="mlp"), seed=0)
    assert all(torch.equal(p, q) for p, q in zip(params, other.parameters(a)))
    with pytest.raises(ValueError):
        mappo.MAPPOLearner(_FakeEnv(), mappo.MAPPOConfig(rollouts=20, sequence_length=16))

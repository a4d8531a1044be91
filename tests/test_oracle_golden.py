"""Pin the CPU oracle (test infrastructure) against the analytic known-answer vectors and against
numpy's own float16 arithmetic — the only anchors available: the reference ships no tests and
pymunk cannot be installed here (PARITY UNPINNED)."""
import numpy as np
import pytest

import parity_utils as pu
from as_cops_and_thieves_b200.maps import compile_map
from oracle import cat_oracle as co

M, VECTORS = pu.load_analytic()
CMAP = compile_map(M, name="analytic")


@pytest.mark.parametrize("vec", VECTORS, ids=[v["name"] for v in VECTORS])
def test_oracle_matches_analytic_vector(vec):
    pu.check_vector(vec, pu.eval_vector_oracle(CMAP, vec))


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32-10
    assert [hex(x) for x in co.philox([0, 0, 0, 0], [0, 0])] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
    assert [hex(x) for x in co.philox([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2)] == \
        ["0x408f276d", "0x41c83b0e", "0xa20bc7c6", "0x6d5451fd"]
    assert [hex(x) for x in co.philox([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0])] == \
        ["0xd16cfe09", "0x94fdcceb", "0x5001e420", "0x24126ea1"]


def test_half_conversions_match_numpy():
    rng = np.random.default_rng(1)
    xs = np.concatenate([rng.uniform(-2000, 2000, 20000), rng.uniform(-1, 1, 5000) * 1e-4,
                         [0.0, 65504.0, 65519.9, 65520.0, 1e-8, 2047.5, 2048.5, 1023.75, 0.1]])
    L = co.lib()
    got = np.array([L.orc_double_to_half_bits(float(x)) for x in xs], np.uint16)
    with np.errstate(over="ignore"):
        want = xs.astype(np.float16).view(np.uint16)
    assert np.array_equal(got, want)
    back = np.array([L.orc_half_bits_to_float(int(b)) for b in got[:1000]], np.float32)
    assert np.array_equal(back, got[:1000].view(np.float16).astype(np.float32))


def test_oracle_f16_chain_equals_numpy_on_its_own_hit_points():
    """entity.py:206-210 executed by numpy on the oracle's fp64 hit points == oracle's f16 output."""
    cm = pu.named_cmap("squarinth")
    orc = co.Oracle(cm, seed=5)
    st = orc.new_state(64)
    out = orc.reset(st)
    pts = out.hit_point.astype(np.float16)                      # np.array(points, dtype=float16) from fp64
    o16 = st.pos.astype(np.float16)
    dx = pts[..., 0] - o16[:, :, None, 0]
    dy = pts[..., 1] - o16[:, :, None, 1]
    chain = np.hypot(dx, dy).astype(np.float16)
    chain = np.where(out.obs_type == pu.TYPE_EMPTY, np.float16(400.0), chain)
    assert np.array_equal(chain.view(np.uint16), out.obs_dist.view(np.uint16))


def test_shared_merge_and_rewards_match_independent_numpy():
    cm = pu.named_cmap("lbirinth")
    orc = co.Oracle(cm, seed=2, auto_reset=0)
    st = orc.new_state(128)
    orc.reset(st)
    rng = np.random.default_rng(3)
    for _ in range(30):
        out = orc.step(st, rng.integers(0, 4, (128, 3)))
    sd, stp = pu.shared_merge_numpy(out.obs_dist, out.obs_type, cm.n_cops)
    assert np.array_equal(stp, out.shared_type)
    assert np.array_equal(sd.view(np.uint16), out.shared_dist.view(np.uint16))
    cap = (out.winner == 0)
    tmo = out.truncated.astype(bool)
    exp = pu.expected_rewards_numpy(out.obs_dist, out.obs_type, cm.n_cops, cap, tmo)
    np.testing.assert_allclose(out.reward, exp, atol=1e-6)
    assert (out.obs_type == pu.TYPE_COP).any() or (out.obs_type == pu.TYPE_THIEF).any()


def test_step_ordering_observation_is_pre_physics():
    """step() observes and rewards the PRE-physics state; positions move afterwards."""
    orc = co.Oracle(CMAP, auto_reset=0)
    st = orc.new_state(1)
    st.pos[0] = [[50.0, 150.0], [900, 700], [1000, 700]]
    st.tc[0] = st.pos[0]
    st.vel[0, 0] = [60.0, 0.0]
    out = orc.step(st, np.array([[2, 1, 1]], np.int32))
    assert float(out.obs_dist[0, 0, 0]) == 49.0                 # from x=50, not from the post-step x
    assert st.pos[0, 0, 0] == pytest.approx(50.0 + 70.0 / 60.0)


def test_stale_shape_cache_after_reset_A10():
    """pymunk keeps the cached shape centre until the next space.step: after reset() rays still see
    the other agents where they were; the sane variant sees the true positions."""
    for stale in (1, 0):
        orc = co.Oracle(CMAP, stale_shape_cache=stale, seed=9)
        st = orc.new_state(1)
        st.pos[0] = [[50.0, 50.0], [80.0, 50.0], [1000.0, 700.0]]
        st.tc[0] = st.pos[0]
        before_tc = st.tc.copy()
        orc.reset(st)                                           # cop_0 and thief_0 re-spawn, cop_1 has no region
        assert np.array_equal(st.pos[0, 1], [950.0, 700.0])     # Entity.reset() -> its initial position
        if stale:
            assert np.array_equal(st.tc, before_tc)
        else:
            assert np.array_equal(st.tc, st.pos)
        out = orc.step(st, np.array([[1, 1, 1]], np.int32))
        assert np.array_equal(st.tc, st.pos)                    # space.step refreshes the caches
        assert out.obs_dist.shape == (1, 3, 90)


def test_spawn_respects_regions_and_rejection():
    cm = pu.named_cmap("squarinth")
    orc = co.Oracle(cm, seed=11, stale_shape_cache=0)
    st = orc.new_state(512)
    orc.reset(st)
    assert np.all((st.pos[:, :2] >= 300) & (st.pos[:, :2] <= 500))   # cops: one 200x200 region
    corners = [(110, 110), (640, 110), (110, 640), (640, 640)]
    tp = st.pos[:, 2]
    in_any = np.zeros(len(tp), bool)
    used = set()
    for k, (x, y) in enumerate(corners):
        inside = (tp[:, 0] >= x) & (tp[:, 0] <= x + 50) & (tp[:, 1] >= y) & (tp[:, 1] <= y + 50)
        in_any |= inside
        if inside.any():
            used.add(k)
    assert in_any.all() and used == {0, 1, 2, 3}
    assert np.all(st.vel == 0) and np.all(st.step_count == 0) and np.all(st.episode == 1)
    # rejection rule: nothing spawns within 6 of a hull
    for w in range(0, 512, 37):
        for a in range(3):
            assert min(orc.hull_distance(h, st.pos[w, a]) for h in range(cm.n_hulls)) >= 6.0 - 1e-4


def test_spawn_is_a_function_of_global_world_id():
    cm = pu.named_cmap("squarinth")
    orc = co.Oracle(cm, seed=4)
    full = orc.new_state(64)
    orc.reset(full)
    lo, hi = orc.new_state(40, gid0=0), orc.new_state(24, gid0=40)
    orc.reset(lo)
    orc.reset(hi)
    assert np.array_equal(np.concatenate([lo.pos, hi.pos]), full.pos)
    other = co.Oracle(cm, seed=5)
    st2 = other.new_state(64)
    other.reset(st2)
    assert not np.array_equal(st2.pos, full.pos)


def test_persistent_contact_warm_start_and_expiry():
    """cpArbiter caching: jnAcc persists while the pair keeps touching, ages out after 3 idle steps."""
    orc = co.Oracle(CMAP, auto_reset=0)
    st = orc.new_state(1)
    st.pos[0] = [[94.2, 150.0], [900, 700], [1000, 700]]
    st.tc[0] = st.pos[0]
    push = np.array([[2, 1, 1]], np.int32)    # cop_0 keeps pushing +x into wall A
    orc.step(st, push)
    assert st.wall_age[0, 0, 0] == 0 and st.wall_jn[0, 0, 0] == pytest.approx(10.0)
    orc.step(st, push)
    assert st.wall_jn[0, 0, 0] == pytest.approx(10.0) and abs(st.vel[0, 0, 0]) < 1e-12
    away = np.array([[0, 1, 1]], np.int32)    # move away: contact persists while d <= 6, then ages out
    ages = []
    for _ in range(40):
        orc.step(st, away)
        ages.append(int(st.wall_age[0, 0, 0]))
    assert -1 in ages
    k = ages.index(-1)
    assert ages[k - 3:k] == [0, 1, 2] or ages[k - 2:k] == [1, 2]


def test_gae_matches_skrl_formula_in_float64():
    rng = np.random.default_rng(0)
    T, M = 37, 23
    r = rng.normal(size=(T, M)).astype(np.float32)
    v = rng.normal(size=(T, M)).astype(np.float32)
    d = rng.random((T, M)) < 0.1
    lv = rng.normal(size=M).astype(np.float32)
    ret, adv = co.gae(r, d, v, lv, 0.99, 0.95, normalize=True)
    a = np.zeros(M)
    advs = np.zeros((T, M))
    for t in reversed(range(T)):
        nv = v[t + 1] if t < T - 1 else lv
        a = r[t] - v[t] + 0.99 * (~d[t]) * (nv + 0.95 * a)
        advs[t] = a
    np.testing.assert_allclose(ret, advs + v, rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(adv, (advs - advs.mean()) / (advs.std(ddof=1) + 1e-8), rtol=1e-5, atol=1e-6)

"""Shared helpers for the CUDA-vs-oracle parity tests (test infrastructure)."""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

from as_cops_and_thieves_b200.maps import (compile_map, free_space_regions, load_named_map)  # noqa: E402
from oracle.cat_oracle import Oracle, OracleState  # noqa: E402

TYPE_WALL, TYPE_COP, TYPE_THIEF, TYPE_EMPTY = 0, 1, 2, 4

#: tolerances from BASELINE.json north_star / BASELINE.md §5
POS_RTOL = 1e-4          # positions / velocities, relative (fp32 vs fp64)
RAY_ATOL = 0.04          # ray hit distance: 1e-4 of the 400 sensor range, before f16 quantisation
REWARD_ATOL = 2e-3       # rewards: fp32 from the f16 distance, one f16 ulp of slack


def named_cmap(name: str, free_spawn: bool = False, **kw):
    m = load_named_map(name)
    so = free_space_regions(m) if free_spawn else None
    return compile_map(m, name=name, spawn_override=so, **kw)


def f16_chain_numpy(hit_xy: np.ndarray, origin_xy: np.ndarray) -> np.ndarray:
    """entity.py:206-210 executed by numpy itself on fp32 hit points (what the reference runs).

    hit_xy [..., R, 2] float32, origin_xy [..., 2] float32 -> float16 distances [..., R].
    The python-float origin of the reference is cast to f16 by numpy's scalar promotion, which is
    what ``origin.astype(np.float16)`` reproduces.
    """
    pts = hit_xy.astype(np.float16)
    o = origin_xy.astype(np.float16)
    dx = pts[..., 0] - o[..., None, 0]
    dy = pts[..., 1] - o[..., None, 1]
    return np.hypot(dx, dy).astype(np.float16)


def cuda_state_to_oracle(orc: Oracle, st: dict, gid0: int = 0) -> OracleState:
    """Convert ``CatWorlds.get_state()`` tensors into the oracle's dense fp64 state."""
    pos = st["pos"].cpu().numpy()
    N, A = pos.shape[0], pos.shape[1]
    o = OracleState(N, A, orc.H, gid0)
    o.pos[...] = pos.astype(np.float64)
    o.vel[...] = st["vel"].cpu().numpy().astype(np.float64)
    o.vbias[...] = st["vbias"].cpu().numpy().astype(np.float64)
    o.tc[...] = st["tc"].cpu().numpy().astype(np.float64)
    o.step_count[...] = st["step_count"].cpu().numpy()
    o.episode[...] = st["episode"].cpu().numpy().astype(np.uint32)
    wh = st["wall_hull"].cpu().numpy()
    wa = st["wall_age"].cpu().numpy()
    wj = st["wall_jn"].cpu().numpy()
    n_idx, a_idx, k_idx = np.nonzero(wh >= 0)
    o.wall_jn[n_idx, a_idx, wh[n_idx, a_idx, k_idx]] = wj[n_idx, a_idx, k_idx]
    o.wall_age[n_idx, a_idx, wh[n_idx, a_idx, k_idx]] = wa[n_idx, a_idx, k_idx]
    pa = st["pair_age"].cpu().numpy()
    pj = st["pair_jn"].cpu().numpy()
    p = 0
    for i in range(A):
        for j in range(i + 1, A):
            o.pair_age[:, i, j] = pa[:, p]
            o.pair_jn[:, i, j] = np.where(pa[:, p] >= 0, pj[:, p], 0.0)
            p += 1
    return o


def perturbed(o: OracleState, rng: np.random.Generator, eps: float) -> OracleState:
    q = o.copy()
    d = rng.choice([-eps, eps], size=q.pos.shape)
    q.pos += d
    moved = np.all(o.tc == o.pos, axis=(1, 2))
    q.tc[moved] += d[moved]                     # centres that coincide with the body move with it
    q.tc[~moved] += rng.choice([-eps, eps], size=q.tc[~moved].shape)
    return q


def ray_unstable_mask(orc: Oracle, base: OracleState, base_alpha: np.ndarray, base_type: np.ndarray,
                      n_pert: int = 4, eps: float = 1e-3, tol: float = 0.01, seed: int = 0) -> np.ndarray:
    """ε-boundary class for rays: a ray is flagged when nudging every body by
    ±eps changes what it hits or moves its hit distance by more than ``tol`` — i.e. tangent /
    corner / threshold / tie cases whose answer is not determined at fp32 resolution."""
    rng = np.random.default_rng(seed)
    L = orc.params["ray_length"]
    bad = np.zeros(base_type.shape, bool)
    for _ in range(n_pert):
        q = perturbed(base, rng, eps)
        out = orc.observe(q)
        bad |= out.obs_type != base_type
        bad |= np.abs(out.hit_alpha - base_alpha) * L > tol
    return bad


def expected_rewards_numpy(obs_dist: np.ndarray, obs_type: np.ndarray, n_cops: int, captured: np.ndarray,
                           timeout: np.ndarray) -> np.ndarray:
    """cop.py:49-75 / thief.py:48-69 evaluated with numpy in float64 from f16 observations."""
    N, A, _ = obs_dist.shape
    out = np.zeros((N, A), np.float64)
    d = obs_dist.astype(np.float64)
    for a in range(A):
        is_cop = a < n_cops
        want = TYPE_THIEF if is_cop else TYPE_COP
        m = obs_type[:, a] == want
        seen = m.any(axis=1)
        dmin = np.where(m, d[:, a], np.inf).min(axis=1)
        if is_cop:
            r = np.where(seen, -0.02 + 1.5 * np.exp(-np.where(seen, dmin, 0.0) / 50.0), -0.04)
            r = np.where(captured, 1.0, np.where(timeout, -1.0, r))
        else:
            r = np.where(seen, np.tanh((np.where(seen, dmin, 0.0) - 100.0) / 50.0) / 10.0, 0.15)
            r = np.where(captured, -1.0, np.where(timeout, 1.0, r))
        out[:, a] = r
    return out


def shared_merge_numpy(obs_dist: np.ndarray, obs_type: np.ndarray, n_cops: int):
    """observation_spaces.py:97-121 net effect computed independently with numpy."""
    N, A, R = obs_type.shape
    sd = np.zeros((N, 2, R), np.float16)
    stp = np.full((N, 2, R), TYPE_EMPTY, np.uint8)
    for team, (a0, a1) in enumerate(((0, n_cops), (n_cops, A))):
        t = np.full((N, R), TYPE_EMPTY, np.uint8)
        d = np.zeros((N, R), np.float16)
        for a in range(a0, a1):
            take = t == TYPE_EMPTY
            t = np.where(take, obs_type[:, a], t)
            d = np.where(take, obs_dist[:, a], d)
        stp[:, team], sd[:, team] = t, d
    return sd, stp


def flat_state_numpy(obs_dist, obs_type, shared_dist, shared_type, team_pos, n_cops):
    """env.state() flattened the way skrl does, built with numpy."""
    N, A, R = obs_type.shape
    cols = []
    for a in range(A):
        team = 0 if a < n_cops else 1
        a0, a1 = (0, n_cops) if team == 0 else (n_cops, A)
        cols += [shared_dist[:, team].astype(np.float32), shared_type[:, team].astype(np.float32),
                 obs_dist[:, a].astype(np.float32), obs_type[:, a].astype(np.float32),
                 team_pos[:, a0:a1].astype(np.float32).reshape(N, -1)]
    return np.concatenate(cols, axis=1)


# ------------------------------------------------------------------ analytic golden vectors
GOLDEN_DIR = ROOT / "tests" / "golden"


def load_analytic():
    import json
    from as_cops_and_thieves_b200.maps import Map
    data = json.load(open(GOLDEN_DIR / "analytic_vectors.json"))
    m = Map(GOLDEN_DIR / "analytic_map.json")
    return m, data["vectors"]


def vector_params(vec) -> dict:
    p = dict(auto_reset=int(vec.get("auto_reset", 0)), max_step_count=int(vec.get("max_step_count", 400)))
    if "capture_radius_override" in vec:
        p["termination_radius"] = float(vec["capture_radius_override"])
    return p


def eval_vector_oracle(cmap, vec) -> dict:
    """Run one analytic vector through the CPU oracle."""
    orc = Oracle(cmap, **vector_params(vec))
    st = orc.new_state(1)
    st.pos[0] = np.asarray(vec["pos"], np.float64)
    st.tc[0] = st.pos[0]
    if "vel" in vec:
        st.vel[0] = np.asarray(vec["vel"], np.float64)
    st.step_count[0] = int(vec.get("step_count", 0))
    if vec["kind"] == "ray":
        out = orc.observe(st)
    else:
        out = orc.step(st, np.asarray([vec["actions"]], np.int32))
    return dict(obs_type=out.obs_type[0], obs_dist=out.obs_dist[0], hit_point=out.hit_point[0],
                reward=out.reward[0], terminated=int(out.terminated[0]), truncated=int(out.truncated[0]),
                winner=int(out.winner[0]), pos=st.pos[0], vel=st.vel[0], vbias=st.vbias[0])


def check_vector(vec, res, pos_tol=1e-6, point_tol=1e-6) -> None:
    """Assert a backend's result against the analytic expectation (tolerances: fp64 oracle by
    default; the CUDA tests pass the fp32 tolerances)."""
    exp = vec["expect"]
    name = vec["name"]
    if vec["kind"] == "ray":
        a, r = vec["agent"], vec["ray"]
        assert int(res["obs_type"][a, r]) == exp["type"], name
        want16 = np.float16(exp["distance"])
        got16 = np.float16(res["obs_dist"][a, r])
        # the f16 chain quantises the hit point and the origin first: allow one f16 step of the analytic value
        step16 = float(np.spacing(np.float16(max(abs(exp["distance"]), 1.0))))
        assert abs(float(got16) - float(want16)) <= 2 * step16, (name, got16, want16)
        if "point" in exp:
            np.testing.assert_allclose(res["hit_point"][a, r], exp["point"], atol=point_tol, err_msg=name)
        return
    for key in ("terminated", "truncated", "winner"):
        if key in exp:
            assert res[key] == exp[key], (name, key, res[key], exp[key])
    if "reward" in exp:
        np.testing.assert_allclose(res["reward"], exp["reward"], atol=REWARD_ATOL, err_msg=name)
    agents = vec.get("check_agents", list(range(len(vec["pos"]))))
    for key in ("pos", "vel", "vbias"):
        if key in exp:
            np.testing.assert_allclose(np.asarray(res[key])[agents], exp[key], atol=pos_tol, err_msg=f"{name}:{key}")


def random_map_json(seed: int, n_blocks: int = 18, size=(900.0, 700.0), n_cops: int = 2, n_thieves: int = 1) -> dict:
    """A random map in the reference's ``maps_templates`` schema: an outer frame of four thin rectangles plus
    rectangles (some with negative extents) and arbitrary — possibly concave, possibly overlapping — polygons, as
    ``agh-map.json`` has them.  Agents spawn in the free margin left along the frame."""
    rng = np.random.default_rng(seed)
    W, H = size
    blocks = [{"type": "rect", "x": 20, "y": 20, "w": W - 40, "h": 6}, {"type": "rect", "x": 20, "y": H - 26, "w": W - 40, "h": 6},
              {"type": "rect", "x": 20, "y": 20, "w": 6, "h": H - 40}, {"type": "rect", "x": W - 26, "y": 20, "w": 6, "h": H - 40}]
    for _ in range(n_blocks):
        cx, cy = rng.uniform(140, W - 140), rng.uniform(140, H - 140)
        if rng.random() < 0.4:
            w, h = rng.uniform(8, 90) * rng.choice([-1, 1]), rng.uniform(8, 90) * rng.choice([-1, 1])
            blocks.append({"type": "rect", "x": float(cx), "y": float(cy), "w": float(w), "h": float(h)})
        else:
            k = int(rng.integers(3, 9))
            ang = np.sort(rng.uniform(0, 2 * np.pi, k))
            rad = rng.uniform(10, 70, k)                     # star-shaped ring: concave in general
            vs = [{"x": float(cx + r * np.cos(a)), "y": float(cy + r * np.sin(a))} for a, r in zip(ang, rad)]
            blocks.append({"type": "poly", "vs": vs})
    agents = []
    for i in range(n_cops + n_thieves):
        y0 = 40 + i * (H - 80) / (n_cops + n_thieves)
        agents.append({"type": "cop" if i < n_cops else "thief", "x": 60.0, "y": float(y0 + 20),
                       "spawn_region": {"x": 36.0, "y": float(y0), "w": 60.0, "h": float((H - 80) / (n_cops + n_thieves) - 10)}})
    return {"window": {"w_px": int(W), "h_px": int(H)}, "canvas": {"w": int(W), "h": int(H)},
            "objects": {"blocks": blocks}, "agents": agents}


def random_cmap(seed: int, tmp_path, **kw):
    import json
    from as_cops_and_thieves_b200.maps import Map
    path = Path(tmp_path) / f"random_{seed}.json"
    path.write_text(json.dumps(random_map_json(seed, **kw)))
    return compile_map(Map(str(path)), name=f"random_{seed}")

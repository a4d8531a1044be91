"""Host-side logic that needs neither a GPU nor the oracle."""
import numpy as np
import pytest

from as_cops_and_thieves_b200 import spaces
from as_cops_and_thieves_b200.env import _agent_observation_space, _shared_observation_space
from as_cops_and_thieves_b200.params import EnvParams, load_physical_params
from as_cops_and_thieves_b200.sharding import shard_range
from as_cops_and_thieves_b200.render import render_rgb
from as_cops_and_thieves_b200.maps import compile_map, load_named_map


def test_observation_space_matches_entity_py():
    sp = _agent_observation_space(90, 400.0)                       # entity.py:92-107
    assert list(sp.spaces.keys()) == ["distance", "object_type"]   # Dict sorts keys -> flatten order
    assert sp["distance"].shape == (90,) and sp["distance"].dtype == np.float16
    assert sp["object_type"].shape == (90,) and sp["object_type"].dtype == np.uint8
    assert float(sp["distance"].high.max()) == 400.0 and int(sp["object_type"].high.max()) == 4
    assert spaces.flatdim(sp) == 180


def test_state_space_layout_is_1090_for_two_cops_one_thief():
    ids = ["cop_0", "cop_1", "thief_0"]
    sh = _shared_observation_space((1280, 800), ids, 2, 90, 400.0)  # observation_spaces.py:13-64
    assert list(sh.spaces.keys()) == ids
    keys = list(sh["cop_0"].spaces.keys())
    assert keys == ["distance_shared", "object_type_shared", "own_distances", "own_obj_types", "team_positions"]
    assert sh["cop_0"]["team_positions"].shape == (2, 2) and sh["thief_0"]["team_positions"].shape == (1, 2)
    assert float(sh["cop_0"]["team_positions"].high.max()) == 1280.0
    assert [spaces.flatdim(sh[a]) for a in ids] == [364, 364, 362]
    assert spaces.flatdim(sh) == 1090


def test_flatten_follows_sorted_key_order():
    sp = _agent_observation_space(4, 400.0)
    x = {"object_type": np.array([4, 0, 1, 2], np.uint8), "distance": np.array([400, 1, 2, 3], np.float16)}
    assert spaces.flatten(sp, x).tolist() == [400, 1, 2, 3, 4, 0, 1, 2]
    assert spaces.flatten(sp, x).dtype == np.float32


def test_discrete_action_space():
    a = spaces.Discrete(4)
    assert a.n == 4 and all(a.contains(i) for i in range(4)) and not a.contains(4)


def test_default_params_are_the_reference_constants():
    p = EnvParams()
    assert (p.unit_velocity, p.unit_mass, p.unit_size, p.max_speed, p.termination_radius) == (10.0, 1.0, 5.0, 125.0, 20.0)
    assert (p.ray_length, p.n_rays, p.ray_radius, p.wall_radius) == (400.0, 90, 1.0, 1.0)
    assert p.dt == pytest.approx(1 / 60) and p.max_step_count == 400
    assert 1.0 - p.collision_bias ** p.dt == pytest.approx(0.1)     # biasCoef at dt = 1/60


def test_physical_params_are_read_from_pyproject(tmp_path):
    f = tmp_path / "pyproject.toml"
    f.write_text("[tool.physical-params]\nunit_velocity = 7.5\nmax_speed = 50\nunit_size = 5.0\n"
                 "unit_mass = 2.0\ntermination_radius = 11.0\npymunk_cop_category = 42\n")
    got = load_physical_params(str(f))
    assert got == dict(unit_velocity=7.5, unit_mass=2.0, unit_size=5.0, max_speed=50.0, termination_radius=11.0)
    assert load_physical_params(str(tmp_path / "missing.toml"))["max_speed"] == 125.0


@pytest.mark.parametrize("n,ws", [(65536, 8), (16384, 3), (10, 4), (7, 8), (1, 1)])
def test_shard_ranges_partition_the_world_ids(n, ws):
    spans = [shard_range(n, r, ws) for r in range(ws)]
    assert spans[0][0] == 0
    for (g0, c0), (g1, _) in zip(spans, spans[1:]):
        assert g0 + c0 == g1
    assert spans[-1][0] + spans[-1][1] == n
    sizes = [c for _, c in spans]
    assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(n, ws, ws)


def test_renderer_draws_walls_and_agents():
    cm = compile_map(load_named_map("squarinth"))
    img = render_rgb(cm, np.array([[350.0, 350.0], [300.0, 300.0], [450.0, 330.0]]))
    assert img.shape == (1280, 800, 3) and img.dtype == np.uint8
    assert tuple(img[102, 400]) == (255, 255, 255)     # inside the left wall x in [100,105]
    assert tuple(img[350, 350]) == (0, 0, 255)         # cop blue
    assert tuple(img[450, 330]) == (255, 0, 0)         # thief red
    assert tuple(img[600, 400]) == (0, 0, 0)


def test_packed_record_types_unpack_to_object_type_values():
    """include/cat_b200.h CatRecordLayout, type_bits == 2: ray r of a world is bits 2 (r % 4) .. + 1 of byte r // 4,
    codes wall 0 / cop 1 / thief 2 / empty 3 (= ObjectType EMPTY, 4).  The decoder of the host-facing path against a
    plain numpy packing, on a ray count that is not a multiple of four."""
    import numpy as np
    import torch
    from as_cops_and_thieves_b200.worlds import CatWorlds, HostRecords
    rng = np.random.default_rng(0)
    N, A, R = 5, 3, 37
    types = rng.choice(np.array([0, 1, 2, 4], np.uint8), size=(N, A * R))
    codes = np.where(types == 4, 3, types).astype(np.uint8)
    padded = np.zeros((N, (A * R + 3) // 4 * 4), np.uint8)
    padded[:, :A * R] = codes
    q = padded.reshape(N, -1, 4)
    packed = (q[..., 0] | (q[..., 1] << 2) | (q[..., 2] << 4) | (q[..., 3] << 6)).astype(np.uint8)
    got = CatWorlds.unpack_types(torch.from_numpy(packed), A, R)
    assert got.shape == (N, A, R) and got.dtype == torch.uint8
    assert np.array_equal(got.numpy().reshape(N, -1), types)
    h = HostRecords(obs_type_packed=torch.from_numpy(packed))
    h.shape = (A, R)
    assert np.array_equal(h["obs_type"].numpy().reshape(N, -1), types)      # decoded on request, not stored
    assert "obs_type" not in h
    with pytest.raises(KeyError):
        h["no-such-view"]

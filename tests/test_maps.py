"""Map loader / compiler (host logic, CPU)."""
import json
import math
from pathlib import Path

import numpy as np
import pytest

from as_cops_and_thieves_b200.maps import (Map, compile_map, convex_hull_ccw, free_space_regions, load_named_map,
                                           builtin_map_path)

GOLDEN = Path(__file__).resolve().parent / "golden"

# SURVEY.md Appendix B (computed from the reference's maps_templates with a monotone-chain hull)
EXPECTED = {"squarinth": (4, 16, 4), "lbirinth": (6, 24, 4), "grandbyrinth": (4, 16, 4),
            "labyrinth": (30, 119, 4), "agh-map": (105, 496, 10)}


@pytest.mark.parametrize("name", list(EXPECTED))
def test_hull_and_edge_counts(name):
    cm = compile_map(load_named_map(name))
    H, E, mx = EXPECTED[name]
    assert (cm.n_hulls, cm.n_edges, int(np.diff(cm.hull_off).max())) == (H, E, mx)
    assert cm.n_cops == 2 and cm.n_thieves == 1
    assert cm.agent_ids == ["cop_0", "cop_1", "thief_0"]


@pytest.mark.parametrize("name", list(EXPECTED))
def test_hulls_are_convex_ccw_with_outward_unit_normals(name):
    cm = compile_map(load_named_map(name))
    for h in range(cm.n_hulls):
        o, e = cm.hull_off[h], cm.hull_off[h + 1]
        v = cm.vert[o:e]
        n = len(v)
        for i in range(n):
            a, b, c = v[i - 1], v[i], v[(i + 1) % n]
            cross = (b[0] - a[0]) * (c[1] - b[1]) - (b[1] - a[1]) * (c[0] - b[0])
            assert cross > 0, "strictly convex, counter-clockwise, no collinear points"
        centroid = v.mean(axis=0)
        for i in range(n):
            nx, ny = cm.normal[o + i]
            assert abs(math.hypot(nx, ny) - 1) < 1e-12
            assert (centroid - v[i]) @ cm.normal[o + i] < 0, "normal points away from the interior"
            edge = v[i] - v[i - 1]
            assert abs(edge @ cm.normal[o + i]) < 1e-9
            assert abs(np.hypot(*edge) - cm.edge_len[o + i]) < 1e-12


def test_rect_rule_defaults_and_negative_extents(tmp_path):
    # map.py:37-52: w/h default to 1; negative extents allowed; ring is closed
    data = {"window": {"w_px": 100, "h_px": 80}, "canvas": {"w": 10, "h": 8},
            "objects": {"blocks": [{"x": 3, "y": 4}, {"type": "rect", "x": 10, "y": 10, "w": -4, "h": 2},
                                   {"type": "poly", "vs": [{"x": 0, "y": 0}, {"x": 2, "y": 0}, {"x": 1, "y": 1},
                                                           {"x": 2, "y": 2}, {"x": 0, "y": 2}]}]},
            "agents": [{"type": "cop", "x": 1, "y": 1}, {"type": "thief", "x": 2, "y": 2}]}
    p = tmp_path / "m.json"
    p.write_text(json.dumps(data))
    m = Map(p)
    assert m.window_dimensions == (100, 80) and m.canvas_dimensions == (10, 8)
    assert m.blocks[0] == [(3, 4), (4, 4), (4, 5), (3, 5), (3, 4)]
    assert m.blocks[1][2] == (6, 12)
    cm = compile_map(m)
    assert list(np.diff(cm.hull_off)) == [4, 4, 4]          # the concave poly is convexified to its hull
    assert [6, 10, 10, 12] in cm.hull_bb.tolist()            # hulls are stored in Morton order, not file order
    assert m.cops_positions == [(1, 1)] and m.thieves_positions == [(2, 2)]
    assert m.agent_spawn_regions == {}


def test_reference_schema_fixture_and_spawn_region_forms():
    m = Map(GOLDEN / "analytic_map.json")
    # agent ids count per type in file order (map.py:76-108); env order is cops then thieves
    assert set(m.agent_spawn_regions) == {"cop_0", "thief_0"}
    assert len(m.agent_spawn_regions["cop_0"]) == 1          # singular spawn_region -> list of one
    assert len(m.agent_spawn_regions["thief_0"]) == 2        # plural spawn_regions
    assert m.cops_positions == [(900, 700), (950, 700)] and m.thieves_positions == [(1000, 700)]
    cm = compile_map(m)
    assert cm.region_off.tolist() == [0, 1, 1, 3]            # cop_0, cop_1 (none), thief_0
    assert list(np.diff(cm.hull_off)) == [4, 4, 4]           # collinear vertex (605,610) dropped


def test_labyrinth_without_agents_raises_like_the_reference():
    with pytest.raises(KeyError):
        Map(builtin_map_path("labyrinth"))                   # map.py:75 KeyError: 'agents'
    m = load_named_map("labyrinth")                          # scaled + injected agents
    assert m.cops_count == 2 and m.thieves_count == 1
    xs = [v[0] for ring in m.blocks for v in ring]
    assert max(xs) > 1000


def test_unknown_block_type_and_missing_xy():
    from as_cops_and_thieves_b200.maps import _parse_block
    with pytest.raises(ValueError):
        _parse_block({"type": "circle"})
    with pytest.raises(ValueError):
        _parse_block({"type": "rect", "x": 1})
    with pytest.raises(ValueError):
        _parse_block({"type": "poly"})


def test_convex_hull_drops_duplicates_and_collinear():
    h = convex_hull_ccw([(0, 0), (1, 0), (2, 0), (2, 2), (0, 2), (0, 0), (1, 1)])
    assert h.tolist() == [[0, 0], [2, 0], [2, 2], [0, 2]]


@pytest.mark.parametrize("name", ["squarinth", "labyrinth", "agh-map"])
def test_grid_lists_are_conservative(name):
    """Every hull within contact reach of a point must be listed in the point's cell."""
    from oracle.cat_oracle import Oracle
    cm = compile_map(load_named_map(name))
    orc = Oracle(cm)
    rng = np.random.default_rng(0)
    lo = np.array([cm.grid_x0, cm.grid_y0])
    hi = lo + np.array([cm.nx, cm.ny]) * cm.cell
    pts = rng.uniform(lo, hi, size=(1500, 2))
    for p in pts:
        c = int((p[1] - cm.grid_y0) // cm.cell) * cm.nx + int((p[0] - cm.grid_x0) // cm.cell)
        con = set(cm.con_cell_hulls[cm.con_cell_off[c]:cm.con_cell_off[c + 1]].tolist())
        for h in range(cm.n_hulls):
            d = orc.hull_distance(h, p)
            if d <= 6.0:
                assert h in con
    # lists are sorted ascending (fixes the arbiter order)
    for c in range(cm.nx * cm.ny):
        seg = cm.con_cell_hulls[cm.con_cell_off[c]:cm.con_cell_off[c + 1]]
        assert np.all(np.diff(seg) > 0)
    assert np.array_equal(cm.edge_hull, np.repeat(np.arange(cm.n_hulls), np.diff(cm.hull_off)))


def test_points_outside_grid_are_far_from_every_hull():
    cm = compile_map(load_named_map("agh-map"))
    assert cm.grid_x0 <= cm.hull_bb[:, 0].min() - 6.0 and cm.grid_y0 <= cm.hull_bb[:, 1].min() - 6.0
    assert cm.grid_x0 + cm.nx * cm.cell >= cm.hull_bb[:, 2].max() + 6.0
    assert cm.grid_y0 + cm.ny * cm.cell >= cm.hull_bb[:, 3].max() + 6.0


def test_free_space_regions_override():
    m = load_named_map("agh-map")
    cm = compile_map(m, spawn_override=free_space_regions(m))
    assert cm.region_off.tolist() == [0, 1, 2, 3]
    assert np.all(cm.regions[:, 2] > 500)


@pytest.mark.parametrize("name", ["squarinth", "lbirinth", "labyrinth", "agh-map"])
def test_view_lists_are_conservative_and_nearest_first(name):
    """Every edge the kernel's per-origin candidate test can accept (faces the origin through its own plane or
    the next plane of the hull, and lies within the sensor reach) must be in the list of the origin's grid
    cell — the lists only shorten the scan, they may never drop a candidate."""
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parent))
    import parity_utils as pu
    cm = pu.named_cmap(name, free_spawn=name in ("labyrinth", "agh-map"))
    assert len(cm.view_cell_off) == cm.nx * cm.ny + 1 and cm.view_cell_off[-1] == len(cm.view_cell_edges)
    assert cm.view_range >= 400 + 2
    E = cm.n_edges
    prev, nxt = np.zeros(E, int), np.zeros(E, int)
    for h in range(cm.n_hulls):
        o, e = int(cm.hull_off[h]), int(cm.hull_off[h + 1])
        prev[o:e] = np.roll(np.arange(o, e), 1)
        nxt[o:e] = np.roll(np.arange(o, e), -1)
    A, B, n, nn = cm.vert[prev], cm.vert, cm.normal, cm.normal[nxt]
    AB = B - A
    rng = np.random.default_rng(0)
    pts = np.column_stack([rng.uniform(cm.grid_x0, cm.grid_x0 + cm.nx * cm.cell, 3000),
                           rng.uniform(cm.grid_y0, cm.grid_y0 + cm.ny * cm.cell, 3000)])
    worst = 0
    for o in pts:
        cx, cy = int((o[0] - cm.grid_x0) / cm.cell), int((o[1] - cm.grid_y0) / cm.cell)
        lst = cm.view_cell_edges[cm.view_cell_off[cy * cm.nx + cx]:cm.view_cell_off[cy * cm.nx + cx + 1]]
        pd = ((o - B) * n).sum(1)
        pdn = ((o - B) * nn).sum(1)
        t = np.clip(((o - A) * AB).sum(1) / (AB ** 2).sum(1), 0, 1)
        d = np.hypot(*(o - (A + AB * t[:, None])).T)
        cand = np.nonzero(((pd > 0) | (pdn > 0)) & (d < 402.0))[0]
        assert set(cand.tolist()) <= set(lst.tolist()), (name, o)
        worst = max(worst, len(lst))
    assert worst <= E
    if name == "agh-map":       # the point of the lists: far fewer than all 496 edges
        sizes = np.diff(cm.view_cell_off)
        assert sizes.mean() < 0.4 * E


@pytest.mark.parametrize("seed", range(6))
def test_random_maps_compile_to_valid_hulls_and_conservative_lists(seed, tmp_path):
    """Arbitrary user maps, not just the five shipped ones: concave / overlapping polygons and negative-extent
    rectangles must become strictly convex CCW hulls with outward unit normals, and both per-cell list kinds
    (contact hulls, sensor candidate edges) must be conservative."""
    import sys
    sys.path.insert(0, str(Path(__file__).resolve().parent))
    import parity_utils as pu
    from oracle.cat_oracle import Oracle
    cm = pu.random_cmap(seed, tmp_path)
    assert cm.n_hulls >= 10 and cm.n_edges == cm.hull_off[-1]
    for h in range(cm.n_hulls):
        o, e = cm.hull_off[h], cm.hull_off[h + 1]
        v = cm.vert[o:e]
        n = len(v)
        assert n >= 3
        for i in range(n):
            a, b, c = v[i - 1], v[i], v[(i + 1) % n]
            assert (b[0] - a[0]) * (c[1] - b[1]) - (b[1] - a[1]) * (c[0] - b[0]) > 0
            assert abs(np.hypot(*cm.normal[o + i]) - 1) < 1e-12 and (v.mean(axis=0) - v[i]) @ cm.normal[o + i] < 0
    orc = Oracle(cm)
    E = cm.n_edges
    prev, nxt = np.zeros(E, int), np.zeros(E, int)
    for h in range(cm.n_hulls):
        o, e = int(cm.hull_off[h]), int(cm.hull_off[h + 1])
        prev[o:e] = np.roll(np.arange(o, e), 1)
        nxt[o:e] = np.roll(np.arange(o, e), -1)
    A, B, nrm, nn = cm.vert[prev], cm.vert, cm.normal, cm.normal[nxt]
    AB = B - A
    rng = np.random.default_rng(seed)
    lo = np.array([cm.grid_x0, cm.grid_y0])
    for p in rng.uniform(lo, lo + np.array([cm.nx, cm.ny]) * cm.cell, size=(400, 2)):
        c = int((p[1] - cm.grid_y0) // cm.cell) * cm.nx + int((p[0] - cm.grid_x0) // cm.cell)
        con = set(cm.con_cell_hulls[cm.con_cell_off[c]:cm.con_cell_off[c + 1]].tolist())
        for h in range(cm.n_hulls):
            if orc.hull_distance(h, p) <= 6.0:
                assert h in con
        view = set(cm.view_cell_edges[cm.view_cell_off[c]:cm.view_cell_off[c + 1]].tolist())
        t = np.clip(((p - A) * AB).sum(1) / (AB ** 2).sum(1), 0, 1)
        d = np.hypot(*(p - (A + AB * t[:, None])).T)
        cand = np.nonzero(((((p - B) * nrm).sum(1) > 0) | (((p - B) * nn).sum(1) > 0)) & (d < 402.0))[0]
        assert set(cand.tolist()) <= view


def _library_ray_lists(cmap, n_rays, ray_length, rsum, cell):
    """The lists the C++ builder (csrc/ray_lists.h) makes at cat_env_create, unpacked to {(cell, ray): [(bound, edge)]}."""
    import ctypes as C
    import sys
    sys.path.insert(0, str(Path(__file__).resolve().parent))
    from test_abi import _desc
    from as_cops_and_thieves_b200 import _lib
    L = _lib.load()
    md, keep = _desc(cmap)
    grid = (C.c_double * 5)()
    ns, no = C.c_int64(0), C.c_int64(0)
    assert L.cat_ray_lists_host(C.byref(md), n_rays, ray_length, rsum, cell, C.byref(grid), None, C.byref(ns), None, C.byref(no)) == 0
    slots = np.zeros(ns.value, np.uint32)
    ovf = np.zeros(no.value, np.uint32)
    assert L.cat_ray_lists_host(C.byref(md), n_rays, ray_length, rsum, cell, C.byref(grid), slots.ctypes.data_as(C.c_void_p),
                                C.byref(ns), ovf.ctypes.data_as(C.c_void_p), C.byref(no)) == 0
    END, LINK = 0x7F80FFFF, 0x80000000
    lists = {}
    for s in range(len(slots) // 4):
        out, words, guard = [], list(slots[4 * s:4 * s + 4]), 0
        while words:
            w = int(words.pop(0))
            if w == END:
                break
            if w & LINK:
                o = w & 0x7FFFFFFF
                assert o % 4 == 0 and o + 4 <= len(ovf)
                words = list(ovf[o:o + 4])
                guard += 1
                assert guard < 1000
                continue
            bound = np.array([w & 0xFFFF0000], np.uint32).view(np.float32)[0]
            out.append((float(bound), w & 0xFFFF))
        lists[s] = out
    return tuple(grid), lists


@pytest.mark.parametrize("name,free,cell", [("squarinth", False, 60.0), ("agh-map", True, 75.0), ("lbirinth", False, 0.0)])
def test_ray_list_builder_matches_its_numpy_statement(name, free, cell):
    """csrc/ray_lists.h (C++, what the kernel walks) against maps.ray_lists (numpy) on the same grid: the same edges per
    (cell, ray) up to the eps band of the inclusion tests, sorted by a lower bound that never exceeds the numpy one, which
    in turn never exceeds the true hit distance of any ray cast from inside the cell."""
    from as_cops_and_thieves_b200.maps import ray_lists
    m = load_named_map(name)
    cmap = compile_map(m, name=name, spawn_override=free_space_regions(m) if free else None)
    R, L, rsum = 90, 400.0, 2.0
    grid, lib = _library_ray_lists(cmap, R, L, rsum, cell)
    x0, y0, c, nx, ny = grid
    assert nx * ny * R == len(lib) and (cell == 0.0 or abs(c - cell) < 1e-6)
    if cell == 0.0:
        assert 30000 <= nx * ny <= 80000 and c >= 6.0           # automatic: about 64k cells, not below 6 units
        return
    tight = ray_lists(cmap, R, L, rsum, eps=0.5e-2, grid=grid, return_bounds=True)
    loose = ray_lists(cmap, R, L, rsum, eps=2e-2, grid=grid, return_bounds=True)

    def as_dict(res):
        off, ent, lb = res
        return off, ent & 0xFFFF, lb
    to, te, tl = as_dict(tight)
    lo, le, ll = as_dict(loose)
    n_entries = 0
    for s in range(len(lib)):
        got = lib[s]
        edges = [e for _, e in got]
        assert len(set(edges)) == len(edges)
        bounds = [b for b, _ in got]
        assert bounds == sorted(bounds)                          # nearest lower bound first
        must = set(te[to[s]:to[s + 1]].tolist())
        may = set(le[lo[s]:lo[s + 1]].tolist())
        assert must <= set(edges) <= may, (s, must - set(edges), set(edges) - may)
        # the bound shrinks as eps grows (the clip region of an edge that runs almost along the ray moves a lot), so
        # the library's (eps 1e-2) lies between the numpy ones for 2e-2 and 0.5e-2, up to its bf16 truncation (downwards)
        hi = dict(zip(te[to[s]:to[s + 1]].tolist(), tl[to[s]:to[s + 1]].tolist()))
        lw = dict(zip(le[lo[s]:lo[s + 1]].tolist(), ll[lo[s]:lo[s + 1]].tolist()))
        for b, e in got:
            if e in hi:
                assert b <= hi[e] + 0.02, (s, e, b, hi[e])
            assert b >= lw[e] * (1 - 1 / 128) - 0.05, (s, e, b, lw[e])
        n_entries += len(got)
    assert n_entries > 0

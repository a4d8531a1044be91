"""Real-Pymunk driver for parity pinning and CPU timing — TEST INFRASTRUCTURE, NOT PRODUCT.

Only ``tests/`` and ``bench.py --impl reference`` / ``cpu_baseline`` import this module.

The arithmetic the reference runs lives in third-party ``pymunk`` (Chipmunk2D), which is not installable in the
build image, so the CPU oracle is a restatement and parity is UNPINNED until this module finds a real
pymunk.  ``probe()`` tries ``import pymunk`` as is and with ``baseline/_ref`` (the location ``.gitignore`` reserves for
a driver-provided reference install) on ``sys.path``.  When it is there:

* ``PymunkWorld`` builds ONE world exactly the way the reference does — ``pymunk.Space()`` (``base_env.py:77``),
  every map block as ``pymunk.Poly(space.static_body, ring, radius=1)`` (``map.py:125-128``), every agent as
  ``pymunk.Body(mass, moment_for_circle(mass, 0, r))`` + ``pymunk.Circle`` with ``ShapeFilter(group, categories)``
  (``entity.py:109-124``; groups counted from 1, cops first, ``base_env.py:90-92``) — and runs the reference's own
  call sequence: line-of-sight capture test (``base_env.py:536-550``), action impulse + clamp (``entity.py:126-134``),
  the 90-ray ``segment_query_first`` sensor (``entity.py:182-198``), ``space.step(dt)`` (``base_env.py:392``).
  Bodies are placed with ``reindex_shapes_for_body`` (SURVEY.md §8c last row) unless the stale-cache quirk is wanted.
* ``ReferenceEnv`` drives the UNMODIFIED reference ``SimpleEnv`` when its whole import chain (pettingzoo, gymnasium,
  shapely, pygame, tomli) is importable from ``baseline/_ref`` or ``/root/reference/src``.
* ``time_reference`` runs one process per host core for the CPU baseline.
"""
from __future__ import annotations

import itertools
import json
import math
import os
import sys
import tempfile
import time
from pathlib import Path
from typing import Optional, Sequence, Tuple

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
REF_INSTALL = ROOT / "baseline" / "_ref"
REF_SOURCE = Path("/root/reference/src")

TYPE_WALL, TYPE_COP, TYPE_THIEF, TYPE_EMPTY = 0, 1, 2, 4
COP_CATEGORY, THIEF_CATEGORY = 42, 2137          # pyproject.toml:17-18


def probe() -> Tuple[Optional[object], str]:
    """(pymunk module or None, reason).  Tries the interpreter's own site-packages, then ``baseline/_ref``."""
    override = os.environ.get("CAT_PYMUNK_PATH")       # e.g. the stand-in used by the harness self-test
    tried = []
    for extra in ([override] if override else []) + [None, str(REF_INSTALL)]:
        if extra is not None:
            if not Path(extra).exists():
                tried.append(f"{extra}: no such directory")
                continue
            if extra not in sys.path:
                sys.path.insert(0, extra)
        try:
            import pymunk  # noqa: PLC0415
            ver = getattr(pymunk, "version", "?")
            chip = getattr(pymunk, "chipmunk_version", "?")
            return pymunk, f"pymunk {ver} (Chipmunk {chip}) from {Path(pymunk.__file__).parent}"
        except ImportError as e:
            tried.append(f"{'site-packages' if extra is None else extra}: {e}")
    return None, "pymunk not importable (" + "; ".join(tried) + ")"


class PymunkWorld:
    """One world in a fresh ``pymunk.Space``, driven through the reference's call sites."""

    def __init__(self, pymunk, blocks: Sequence[Sequence[Tuple[float, float]]], agents: Sequence[Tuple[str, Tuple[float, float]]],
                 *, dt: float = 1 / 60.0, unit_velocity: float = 10.0, unit_mass: float = 1.0, unit_size: float = 5.0,
                 max_speed: float = 125.0, termination_radius: float = 20.0, ray_length: float = 400.0,
                 ray_radius: float = 1.0, wall_radius: float = 1.0, n_rays: int = 90, max_step_count: int = 400):
        self.pm = pymunk
        self.dt, self.speed, self.max_speed = dt, unit_velocity, max_speed
        self.term_r, self.ray_len, self.ray_r, self.n_rays = termination_radius, ray_length, ray_radius, n_rays
        self.max_step_count = max_step_count
        self.step_count = 0
        self.space = pymunk.Space()                                          # base_env.py:77
        for ring in blocks:                                                  # map.py:125-128
            self.space.add(pymunk.Poly(self.space.static_body, [tuple(map(float, v)) for v in ring], radius=wall_radius))
        group = itertools.count(1)                                           # base_env.py:90
        ordered = [a for a in agents if a[0] == "cop"] + [a for a in agents if a[0] == "thief"]   # base_env.py:91-96
        self.kinds = [k for k, _ in ordered]
        self.n_cops = sum(1 for k in self.kinds if k == "cop")
        self.bodies, self.shapes, self.filters, self.cats = [], [], [], []
        for kind, start in ordered:                                          # entity.py:109-124
            cat = COP_CATEGORY if kind == "cop" else THIEF_CATEGORY
            g = next(group)
            body = pymunk.Body(unit_mass, pymunk.moment_for_circle(unit_mass, inner_radius=0.0, outer_radius=unit_size))
            body.position = tuple(map(float, start))
            shape = pymunk.Circle(body, radius=unit_size)
            shape.filter = pymunk.ShapeFilter(group=g, categories=cat)
            self.space.add(body, shape)
            self.bodies.append(body); self.shapes.append(shape); self.cats.append(cat)
            self.filters.append(pymunk.ShapeFilter(group=g, categories=cat))
        ang = np.linspace(0.0, 2.0 * np.pi, n_rays, endpoint=False)          # entity.py:182
        self.cos, self.sin = np.cos(ang), np.sin(ang)
        self.force = {0: pymunk.Vec2d(-unit_velocity, 0), 1: pymunk.Vec2d(0, unit_velocity),
                      2: pymunk.Vec2d(unit_velocity, 0), 3: pymunk.Vec2d(0, -unit_velocity)}    # entity.py:77-82

    # ------------------------------------------------------------------ state
    def set_state(self, pos, vel=None, reindex: bool = True) -> None:
        for a, body in enumerate(self.bodies):
            body.position = (float(pos[a][0]), float(pos[a][1]))            # entity.py:154-156
            if vel is not None:
                body.velocity = (float(vel[a][0]), float(vel[a][1]))
            if reindex:                                                     # the sane variant
                self.space.reindex_shapes_for_body(body)

    def get_state(self):
        return (np.array([[b.position[0], b.position[1]] for b in self.bodies]),
                np.array([[b.velocity[0], b.velocity[1]] for b in self.bodies]))

    # ------------------------------------------------------------------ reference call sites
    def observe(self):
        """entity.py:182-198 for every agent: (alpha [A, R], point [A, R, 2], type [A, R])."""
        A, R, pm = len(self.bodies), self.n_rays, self.pm
        alpha = np.ones((A, R)); point = np.zeros((A, R, 2)); typ = np.full((A, R), TYPE_EMPTY, np.uint8)
        for a, body in enumerate(self.bodies):
            origin = body.position
            ends = np.column_stack((origin[0] + self.ray_len * self.cos, origin[1] + self.ray_len * self.sin))
            for i, end in enumerate(ends):
                hit = self.space.segment_query_first(origin, pm.Vec2d(*end), self.ray_r, self.filters[a])
                if hit is None:
                    point[a, i] = end
                    continue
                alpha[a, i] = hit.alpha
                point[a, i] = (hit.point[0], hit.point[1])
                typ[a, i] = self._classify(hit.shape)                       # entity.py:222-241
        return alpha, point, typ

    def _classify(self, shape) -> int:
        pm = self.pm
        if shape.body.body_type == pm.Body.DYNAMIC and isinstance(shape, pm.Circle):
            return TYPE_THIEF if shape.filter.categories == THIEF_CATEGORY else TYPE_COP
        return TYPE_WALL

    def captured(self) -> bool:
        """base_env.py:536-550: thief-major, walls-only line of sight, dist < radius (strict)."""
        pm = self.pm
        for t in range(self.n_cops, len(self.bodies)):
            for c in range(self.n_cops):
                flt = pm.ShapeFilter(mask=~(self.cats[t] | self.cats[c]) & 0xFFFFFFFF)
                hit = self.space.segment_query_first(self.bodies[t].position, self.bodies[c].position, 0.0, flt)
                if hit is None and self.bodies[t].position.get_distance(self.bodies[c].position) < self.term_r:
                    return True
        return False

    def step(self, actions: Sequence[int]):
        """BaseEnv.step (base_env.py:354-413) on this world: returns the pre-physics observation tuple and flags."""
        self.step_count += 1
        cap = self.captured()
        timeout = (not cap) and self.step_count >= self.max_step_count
        for a, body in enumerate(self.bodies):                              # entity.py:126-134
            body.apply_impulse_at_local_point(self.force[int(actions[a])])
            if abs(body.velocity) > self.max_speed:
                body.velocity = body.velocity.normalized() * self.max_speed
        obs = self.observe()
        self.space.step(self.dt)                                            # base_env.py:392
        return obs, cap, timeout

    def spawn_blocked(self, agent: int, p, radius: Optional[float] = None) -> bool:
        """_get_non_colliding_position's acceptance test (base_env.py:154-158)."""
        r = self.shapes[agent].radius if radius is None else radius
        return self.space.point_query_nearest((float(p[0]), float(p[1])), r, self.filters[agent]) is not None


def world_from_map(pymunk, m, **params) -> PymunkWorld:
    """``m``: as_cops_and_thieves_b200.maps.Map (same JSON, same rings as the reference's shapely polygons)."""
    agents = [(a["type"], (a["x"], a["y"])) for a in m._agents]
    return PymunkWorld(pymunk, m.blocks, agents, **params)


# ---------------------------------------------------------------------------------- the unmodified reference env
def reference_map_json(m) -> dict:
    """A Map in the reference's ``maps_templates`` schema (map.py:63-117), blocks as polygons, agents with regions."""
    w, h = m.window_dimensions
    cw, ch = m.canvas_dimensions
    counts, agents = {}, []
    for a in m._agents:
        idx = counts.get(a["type"], 0)
        counts[a["type"]] = idx + 1
        ent = {"type": a["type"], "x": a["x"], "y": a["y"]}
        regs = m.agent_spawn_regions.get(f"{a['type']}_{idx}")
        if regs:
            ent["spawn_regions"] = regs
        agents.append(ent)
    return {"window": {"w_px": w, "h_px": h}, "canvas": {"w": cw, "h": ch},
            "objects": {"blocks": [{"type": "poly", "vs": [{"x": x, "y": y} for x, y in ring]} for ring in m.blocks]},
            "agents": agents}


def import_reference_env() -> Tuple[Optional[type], Optional[type], str]:
    """(SimpleEnv, Map, reason) of the unmodified reference, or (None, None, why not)."""
    pm, why = probe()
    if pm is None:
        return None, None, why
    errs = []
    for base in (REF_INSTALL, REF_SOURCE):
        if not base.exists():
            errs.append(f"{base}: absent")
            continue
        if str(base) not in sys.path:
            sys.path.insert(0, str(base))
        try:
            from environments.simple_env import SimpleEnv  # noqa: PLC0415
            from maps.map import Map as RefMap  # noqa: PLC0415
            return SimpleEnv, RefMap, f"reference env from {base}"
        except Exception as e:  # ImportError of pettingzoo / gymnasium / shapely / pygame / tomli ...
            errs.append(f"{base}: {type(e).__name__}: {e}")
    return None, None, "reference SimpleEnv not importable (" + "; ".join(errs) + ")"


class ReferenceEnv:
    """The reference's own ``SimpleEnv`` on one of our maps (run from a temp dir holding a ``pyproject.toml`` with the
    ``[tool.physical-params]`` table, because ``toml_utils.py:41-46`` reads it from the CWD on every call)."""

    PYPROJECT = ("[tool.physical-params]\nunit_velocity = 10.0\nunit_mass = 1.0\nunit_size = 5.0\nmax_speed = 125.0\n"
                 "pymunk_cop_category = 42\npymunk_thief_category = 2137\ntermination_radius = 20.0\n")

    def __init__(self, m, max_step_count: int = 400):
        SimpleEnv, RefMap, why = import_reference_env()
        if SimpleEnv is None:
            raise ImportError(why)
        self.tmp = tempfile.mkdtemp(prefix="cat_ref_")
        (Path(self.tmp) / "pyproject.toml").write_text(self.PYPROJECT)
        (Path(self.tmp) / "map.json").write_text(json.dumps(reference_map_json(m)))
        self._cwd = os.getcwd()
        os.chdir(self.tmp)
        os.environ.setdefault("SDL_VIDEODRIVER", "dummy")
        self.env = SimpleEnv(RefMap(str(Path(self.tmp) / "map.json")), render_mode="rgb_array", max_step_count=max_step_count)
        self.entities = list(self.env.cops) + list(self.env.thieves)

    def set_state(self, pos, vel=None) -> None:
        for a, ent in enumerate(self.entities):
            ent.body.position = (float(pos[a][0]), float(pos[a][1]))
            if vel is not None:
                ent.body.velocity = (float(vel[a][0]), float(vel[a][1]))
            self.env.space.reindex_shapes_for_body(ent.body)

    def get_state(self):
        return (np.array([[e.body.position[0], e.body.position[1]] for e in self.entities]),
                np.array([[e.body.velocity[0], e.body.velocity[1]] for e in self.entities]))

    def step(self, actions: Sequence[int]):
        ids = self.env.possible_agents
        self.env.agents = list(ids)
        return self.env.step({aid: int(a) for aid, a in zip(ids, actions)})

    def close(self) -> None:
        os.chdir(self._cwd)


# ---------------------------------------------------------------------------------- CPU baseline timing
def _time_worker(args):
    map_name, free, seconds, seed, use_env = args
    sys.path.insert(0, str(ROOT))
    from as_cops_and_thieves_b200.maps import load_named_map  # noqa: PLC0415
    pm, _ = probe()
    m = load_named_map(map_name)
    rng = np.random.default_rng(seed)
    if use_env:
        ref = ReferenceEnv(m)
        ref.env.reset(seed=seed)
        step = lambda: ref.step(rng.integers(0, 4, len(ref.entities)))   # noqa: E731
        n_agents, reset = len(ref.entities), lambda: ref.env.reset()
    else:
        w = world_from_map(pm, m)
        step = lambda: w.step(rng.integers(0, 4, len(w.bodies)))           # noqa: E731
        n_agents, reset = len(w.bodies), lambda: None
    for _ in range(20):
        step()
    t0, n = time.perf_counter(), 0
    while time.perf_counter() - t0 < seconds:
        for _ in range(20):
            out = step()
            n += 1
            if use_env and not ref.env.agents:
                reset()
    return n * n_agents, time.perf_counter() - t0


def time_reference(map_name: str, free: bool, seconds: float = 10.0, procs: Optional[int] = None) -> Optional[dict]:
    """agent-steps/s of the real Pymunk path, one process per host core; None if no pymunk."""
    pm, why = probe()
    if pm is None:
        return None
    import multiprocessing as mp  # noqa: PLC0415
    procs = procs or os.cpu_count() or 1
    use_env = import_reference_env()[0] is not None
    with mp.get_context("spawn").Pool(procs) as pool:
        res = pool.map(_time_worker, [(map_name, free, seconds, 100 + i, use_env) for i in range(procs)])
    return {"value": sum(n / dt for n, dt in res), "cores": procs, "seconds": max(dt for _, dt in res),
            "how": ("unmodified reference SimpleEnv.step" if use_env else "reference call sequence on pymunk.Space (PymunkWorld)")
            + f", one world per process, random actions; {why}"}

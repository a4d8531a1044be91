"""STAND-IN for the ``pymunk`` API subset the reference touches — harness self-test only.

Real pymunk is not installable in the build image, so ``tests/test_pymunk_parity.py`` would never execute a line of
its comparison code here.  This module lets it run end to end: it implements ``Space`` / ``Body`` / ``Circle`` /
``Poly`` / ``ShapeFilter`` / ``Vec2d`` with the calling conventions of pymunk 6.x and answers the queries with the
CPU oracle's routines.  Agreement with it proves that the harness drives the API and converts states correctly; it
says NOTHING about parity with Chipmunk (it is the oracle talking to itself) — the harness reports which one it ran
against.  Never imported by the product or by the oracle.
"""
from __future__ import annotations

import math
import types

import numpy as np

version = "stand-in (oracle-backed, NOT pymunk)"
chipmunk_version = "none"
IS_STAND_IN = True


class Vec2d(tuple):
    def __new__(cls, x, y=None):
        if y is None:
            x, y = x
        return tuple.__new__(cls, (float(x), float(y)))
    x = property(lambda s: s[0])
    y = property(lambda s: s[1])
    def __add__(s, o): return Vec2d(s[0] + o[0], s[1] + o[1])
    def __sub__(s, o): return Vec2d(s[0] - o[0], s[1] - o[1])
    def __mul__(s, k): return Vec2d(s[0] * k, s[1] * k)
    __rmul__ = __mul__
    def __truediv__(s, k): return Vec2d(s[0] / k, s[1] / k)
    def __abs__(s): return math.hypot(s[0], s[1])
    length = property(lambda s: math.hypot(s[0], s[1]))
    def normalized(s):
        n = abs(s)
        return Vec2d(s[0] / n, s[1] / n) if n else Vec2d(0.0, 0.0)
    def get_distance(s, o): return math.hypot(s[0] - o[0], s[1] - o[1])


class ShapeFilter:
    def __init__(self, group=0, categories=0xFFFFFFFF, mask=0xFFFFFFFF):
        self.group, self.categories, self.mask = group, categories, mask
    @staticmethod
    def ALL_MASKS(): return 0xFFFFFFFF


def moment_for_circle(mass, inner_radius, outer_radius, offset=(0, 0)):
    return mass * (0.5 * (inner_radius ** 2 + outer_radius ** 2) + offset[0] ** 2 + offset[1] ** 2)


class Body:
    DYNAMIC, KINEMATIC, STATIC = 0, 1, 2

    def __init__(self, mass=0.0, moment=0.0, body_type=0):
        self.mass, self.moment, self.body_type = mass, moment, body_type
        self._p, self._v = Vec2d(0, 0), Vec2d(0, 0)
        self.shapes, self.space = [], None
    position = property(lambda s: s._p, lambda s, v: setattr(s, "_p", Vec2d(v[0], v[1])))
    velocity = property(lambda s: s._v, lambda s, v: setattr(s, "_v", Vec2d(v[0], v[1])))

    def apply_impulse_at_local_point(self, impulse, point=(0, 0)):
        self._v = Vec2d(self._v[0] + impulse[0] / self.mass, self._v[1] + impulse[1] / self.mass)


class Shape:
    def __init__(self, body):
        self.body, self.filter, self.color = body, ShapeFilter(), None
        if body is not None:
            body.shapes.append(self)


class Circle(Shape):
    def __init__(self, body, radius, offset=(0, 0)):
        super().__init__(body)
        self.radius = radius
        self._cached = Vec2d(body.position)     # cpShapeCacheBB happens at space.add / step / reindex


class Poly(Shape):
    def __init__(self, body, vertices, transform=None, radius=0):
        super().__init__(body)
        self.vertices, self.radius = [tuple(map(float, v)) for v in vertices], radius


SegmentQueryInfo = types.SimpleNamespace
PointQueryInfo = types.SimpleNamespace


class Space:
    def __init__(self):
        self.static_body = Body(body_type=Body.STATIC)
        self.polys, self.circles = [], []
        self._orc = self._st = None
        self._dt = None

    def add(self, *objs):
        for o in objs:
            if isinstance(o, Poly):
                self.polys.append(o)
            elif isinstance(o, Circle):
                o._cached = Vec2d(o.body.position)
                self.circles.append(o)
            elif isinstance(o, Body):
                o.space = self
        self._orc = None

    # ------------------------------------------------------------------ oracle plumbing
    def _build(self, dt=None):
        if self._orc is not None and (dt is None or dt == self._dt):
            return
        import sys
        from pathlib import Path
        root = Path(__file__).resolve().parents[3]
        if str(root) not in sys.path:
            sys.path.insert(0, str(root))
        from as_cops_and_thieves_b200.maps import convex_hull_ccw
        from oracle.cat_oracle import Oracle
        hulls = [convex_hull_ccw(p.vertices) for p in self.polys]
        off = np.zeros(len(hulls) + 1, np.int32)
        for h, hv in enumerate(hulls):
            off[h + 1] = off[h] + len(hv)
        A = len(self.circles)
        n_cops = sum(1 for c in self.circles if c.filter.categories == 42)
        cm = types.SimpleNamespace(hull_off=off, vert=np.concatenate(hulls), n_hulls=len(hulls), n_edges=int(off[-1]),
                                   n_cops=n_cops, n_thieves=A - n_cops, n_agents=A,
                                   init_pos=np.array([c.body.position for c in self.circles], float),
                                   region_off=np.zeros(A + 1, np.int32), regions=np.zeros((1, 4)))
        self._dt = dt if dt is not None else (self._dt or 1 / 60.0)
        keep = self._st
        self._orc = Oracle(cm, dt=self._dt, wall_radius=self.polys[0].radius if self.polys else 1.0,
                           unit_size=self.circles[0].radius, unit_mass=self.circles[0].body.mass, auto_reset=0)
        self._st = self._orc.new_state(1)
        if keep is not None and keep.wall_jn.shape == self._st.wall_jn.shape:
            for name in ("vbias", "wall_jn", "wall_age", "pair_jn", "pair_age"):
                getattr(self._st, name)[...] = getattr(keep, name)

    def _tc(self):
        return np.array([c._cached for c in self.circles], float)

    def _self_agent(self, flt):
        if flt.mask != 0xFFFFFFFF:            # the capture test masks out both agent categories: walls only
            return -1, True
        for a, c in enumerate(self.circles):
            if flt.group != 0 and c.filter.group == flt.group:
                return a, False
        return -2, False                       # sees every agent

    def _shape(self, sid):
        H = len(self.polys)
        return self.polys[sid] if sid < H else self.circles[sid - H]

    # ------------------------------------------------------------------ pymunk API
    def reindex_shapes_for_body(self, body):
        for s in body.shapes:
            if isinstance(s, Circle):
                s._cached = Vec2d(body.position)

    def segment_query_first(self, start, end, radius, shape_filter):
        self._build()
        me, walls_only = self._self_agent(shape_filter)
        if not walls_only and me < 0:
            raise NotImplementedError("stand-in: a query that sees every agent is not something the reference issues")
        sid, alpha, pt = self._orc.segment_query_first(self._tc(), -1 if walls_only else me, start, end, radius)
        if sid < 0:
            return None
        return SegmentQueryInfo(shape=self._shape(sid), point=Vec2d(pt[0], pt[1]), normal=Vec2d(0, 0), alpha=alpha)

    def point_query_nearest(self, point, max_distance, shape_filter):
        self._build()
        me, _ = self._self_agent(shape_filter)
        sid, d = self._orc.point_query_nearest(self._tc(), me, point, max_distance)
        if sid < 0:
            return None
        return PointQueryInfo(shape=self._shape(sid), point=Vec2d(point[0], point[1]), distance=d, gradient=Vec2d(0, 0))

    def step(self, dt):
        self._build(dt)
        st = self._st
        for a, c in enumerate(self.circles):
            st.pos[0, a] = c.body.position
            st.vel[0, a] = c.body.velocity
        self._orc.space_step(st)
        for a, c in enumerate(self.circles):
            c.body.position = st.pos[0, a]
            c.body.velocity = st.vel[0, a]
            c._cached = Vec2d(st.pos[0, a])

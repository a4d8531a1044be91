"""ctypes binding of the CPU oracle (``oracle/cat_oracle.c``) — TEST INFRASTRUCTURE, NOT PRODUCT.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs import this module.  PARITY UNPINNED: see the header of ``cat_oracle.h``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path
from typing import Optional

import numpy as np

HERE = Path(__file__).resolve().parent
LIB_PATH = HERE / "libcat_oracle.so"


def build(force: bool = False) -> Path:
    """Compile the oracle with the system gcc (OpenMP when libgomp is usable)."""
    src = [HERE / "cat_oracle.c", HERE / "cat_oracle.h", HERE.parent / "include" / "cat_philox.h"]
    if LIB_PATH.exists() and not force and all(LIB_PATH.stat().st_mtime >= s.stat().st_mtime for s in src):
        return LIB_PATH
    base = ["-O2", "-fPIC", "-std=gnu11", "-ffp-contract=off", "-shared", "-o", str(LIB_PATH), str(src[0]), "-lm"]
    last = None
    for cc in ("/usr/bin/gcc", "gcc", "cc"):
        for omp in (["-fopenmp"], []):
            try:
                subprocess.run([cc] + omp + base, check=True, capture_output=True, text=True)
                return LIB_PATH
            except (subprocess.CalledProcessError, FileNotFoundError) as e:  # try next
                last = e
    raise RuntimeError(f"could not build the oracle: {getattr(last, 'stderr', last)}")


class _OrcMap(C.Structure):
    _fields_ = [("n_hulls", C.c_int32), ("n_edges", C.c_int32), ("hull_off", C.c_void_p), ("vert", C.c_void_p),
                ("n_cops", C.c_int32), ("n_thieves", C.c_int32), ("init_pos", C.c_void_p),
                ("region_off", C.c_void_p), ("regions", C.c_void_p)]


class _OrcParams(C.Structure):
    _fields_ = [("dt", C.c_double), ("max_step_count", C.c_int32),
                ("unit_velocity", C.c_double), ("unit_mass", C.c_double), ("unit_size", C.c_double),
                ("max_speed", C.c_double), ("termination_radius", C.c_double),
                ("ray_length", C.c_double), ("ray_radius", C.c_double), ("wall_radius", C.c_double),
                ("n_rays", C.c_int32), ("iterations", C.c_int32),
                ("collision_slop", C.c_double), ("collision_bias", C.c_double),
                ("collision_persistence", C.c_int32), ("stale_shape_cache", C.c_int32),
                ("auto_reset", C.c_int32), ("seed", C.c_uint64)]


class _OrcState(C.Structure):
    _fields_ = [("n_worlds", C.c_int32), ("gid0", C.c_int64),
                ("pos", C.c_void_p), ("vel", C.c_void_p), ("vbias", C.c_void_p), ("tc", C.c_void_p),
                ("step_count", C.c_void_p), ("episode", C.c_void_p),
                ("wall_jn", C.c_void_p), ("wall_age", C.c_void_p), ("pair_jn", C.c_void_p), ("pair_age", C.c_void_p)]


class _OrcOut(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("obs_dist", "obs_type", "hit_point", "hit_alpha", "reward", "terminated",
                                          "truncated", "winner", "shared_dist", "shared_type", "team_pos")]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(str(LIB_PATH))
        _lib.orc_create.restype = C.c_void_p
        _lib.orc_create.argtypes = [C.POINTER(_OrcMap), C.POINTER(_OrcParams)]
        _lib.orc_destroy.argtypes = [C.c_void_p]
        _lib.orc_num_threads.restype = C.c_int
        _lib.orc_init_state.argtypes = [C.c_void_p, C.POINTER(_OrcState)]
        _lib.orc_reset.argtypes = [C.c_void_p, C.POINTER(_OrcState), C.c_void_p, C.POINTER(_OrcOut)]
        _lib.orc_step.argtypes = [C.c_void_p, C.POINTER(_OrcState), C.c_void_p, C.POINTER(_OrcOut)]
        _lib.orc_observe.argtypes = [C.c_void_p, C.POINTER(_OrcState), C.POINTER(_OrcOut)]
        _lib.orc_gae.argtypes = [C.c_void_p] * 6 + [C.c_int, C.c_int, C.c_double, C.c_double, C.c_int]
        _lib.orc_double_to_half_bits.restype = C.c_uint16
        _lib.orc_double_to_half_bits.argtypes = [C.c_double]
        _lib.orc_half_bits_to_float.restype = C.c_float
        _lib.orc_half_bits_to_float.argtypes = [C.c_uint16]
        _lib.orc_segment_query_first.restype = C.c_int
        _lib.orc_segment_query_first.argtypes = [C.c_void_p, C.c_void_p, C.c_int] + [C.c_double] * 5 + [C.c_void_p, C.c_void_p]
        _lib.orc_space_step.argtypes = [C.c_void_p, C.POINTER(_OrcState)]
        _lib.orc_point_query_nearest.restype = C.c_int
        _lib.orc_point_query_nearest.argtypes = [C.c_void_p, C.c_void_p, C.c_int] + [C.c_double] * 3 + [C.c_void_p]
        _lib.orc_hull_distance.restype = C.c_double
        _lib.orc_hull_distance.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double]
        _lib.orc_philox.argtypes = [C.c_uint32] * 6 + [C.c_void_p]
    return _lib


def _p(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


#: defaults = /root/reference/pyproject.toml:12-19 + Chipmunk space defaults +
#: SimpleEnv defaults (/root/reference/src/environments/simple_env.py:14-20)
DEFAULT_PARAMS = dict(
    dt=1.0 / 60.0, max_step_count=400, unit_velocity=10.0, unit_mass=1.0, unit_size=5.0, max_speed=125.0,
    termination_radius=20.0, ray_length=400.0, ray_radius=1.0, wall_radius=1.0, n_rays=90, iterations=10,
    collision_slop=0.1, collision_bias=(1.0 - 0.1) ** 60.0, collision_persistence=3, stale_shape_cache=1,
    auto_reset=1, seed=0,
)


class OracleState:
    """Caller-owned SoA state for N worlds (numpy, fp64)."""

    def __init__(self, n_worlds: int, A: int, H: int, gid0: int = 0):
        self.N, self.A, self.H, self.gid0 = n_worlds, A, H, gid0
        self.pos = np.zeros((n_worlds, A, 2))
        self.vel = np.zeros((n_worlds, A, 2))
        self.vbias = np.zeros((n_worlds, A, 2))
        self.tc = np.zeros((n_worlds, A, 2))
        self.step_count = np.zeros(n_worlds, np.int32)
        self.episode = np.zeros(n_worlds, np.uint32)
        self.wall_jn = np.zeros((n_worlds, A, H))
        self.wall_age = np.full((n_worlds, A, H), -1, np.int8)
        self.pair_jn = np.zeros((n_worlds, A, A))
        self.pair_age = np.full((n_worlds, A, A), -1, np.int8)

    def c_struct(self) -> _OrcState:
        for name in ("pos", "vel", "vbias", "tc", "step_count", "episode", "wall_jn", "wall_age", "pair_jn", "pair_age"):
            a = getattr(self, name)
            assert a.flags["C_CONTIGUOUS"], name
        return _OrcState(self.N, self.gid0, _p(self.pos), _p(self.vel), _p(self.vbias), _p(self.tc),
                         _p(self.step_count), _p(self.episode), _p(self.wall_jn), _p(self.wall_age),
                         _p(self.pair_jn), _p(self.pair_age))

    def copy(self) -> "OracleState":
        s = OracleState(self.N, self.A, self.H, self.gid0)
        for name in ("pos", "vel", "vbias", "tc", "step_count", "episode", "wall_jn", "wall_age", "pair_jn", "pair_age"):
            getattr(s, name)[...] = getattr(self, name)
        return s


class OracleOut:
    def __init__(self, N: int, A: int, R: int):
        self.obs_dist = np.zeros((N, A, R), np.float16)
        self.obs_type = np.zeros((N, A, R), np.uint8)
        self.hit_point = np.zeros((N, A, R, 2))
        self.hit_alpha = np.zeros((N, A, R))
        self.reward = np.zeros((N, A), np.float32)
        self.terminated = np.zeros(N, np.uint8)
        self.truncated = np.zeros(N, np.uint8)
        self.winner = np.zeros(N, np.int8)
        self.shared_dist = np.zeros((N, 2, R), np.float16)
        self.shared_type = np.zeros((N, 2, R), np.uint8)
        self.team_pos = np.zeros((N, A, 2), np.float16)

    def c_struct(self) -> _OrcOut:
        return _OrcOut(*[_p(getattr(self, n)) for n, _ in _OrcOut._fields_])


class Oracle:
    """One map + parameter set.  ``cmap`` is a ``CompiledMap`` (host map compiler output)."""

    def __init__(self, cmap, **params):
        L = lib()
        self.cmap = cmap
        self.params = dict(DEFAULT_PARAMS)
        self.params.update(params)
        self._keep = [np.ascontiguousarray(cmap.hull_off, np.int32), np.ascontiguousarray(cmap.vert, np.float64),
                      np.ascontiguousarray(cmap.init_pos, np.float64), np.ascontiguousarray(cmap.region_off, np.int32),
                      np.ascontiguousarray(cmap.regions, np.float64)]
        m = _OrcMap(cmap.n_hulls, cmap.n_edges, _p(self._keep[0]), _p(self._keep[1]), cmap.n_cops, cmap.n_thieves,
                    _p(self._keep[2]), _p(self._keep[3]), _p(self._keep[4]))
        pr = _OrcParams(**{k: self.params[k] for k, _ in _OrcParams._fields_})
        self.A, self.H, self.R = cmap.n_agents, cmap.n_hulls, int(self.params["n_rays"])
        self._h = L.orc_create(C.byref(m), C.byref(pr))
        if not self._h:
            raise RuntimeError("orc_create failed")

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_destroy(self._h)
            self._h = None

    def new_state(self, n_worlds: int, gid0: int = 0) -> OracleState:
        st = OracleState(n_worlds, self.A, self.H, gid0)
        cs = st.c_struct()
        lib().orc_init_state(self._h, C.byref(cs))
        return st

    def new_out(self, n_worlds: int) -> OracleOut:
        return OracleOut(n_worlds, self.A, self.R)

    def reset(self, st: OracleState, mask: Optional[np.ndarray] = None, out: Optional[OracleOut] = None) -> OracleOut:
        out = out or self.new_out(st.N)
        cs, co = st.c_struct(), out.c_struct()
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        lib().orc_reset(self._h, C.byref(cs), _p(m), C.byref(co))
        return out

    def step(self, st: OracleState, actions: np.ndarray, out: Optional[OracleOut] = None) -> OracleOut:
        out = out or self.new_out(st.N)
        a = np.ascontiguousarray(actions, np.int32)
        assert a.shape == (st.N, self.A)
        cs, co = st.c_struct(), out.c_struct()
        lib().orc_step(self._h, C.byref(cs), _p(a), C.byref(co))
        return out

    def observe(self, st: OracleState, out: Optional[OracleOut] = None) -> OracleOut:
        out = out or self.new_out(st.N)
        cs, co = st.c_struct(), out.c_struct()
        lib().orc_observe(self._h, C.byref(cs), C.byref(co))
        return out

    def segment_query_first(self, tc: np.ndarray, self_agent: int, a, b, radius: float):
        tc = np.ascontiguousarray(tc, np.float64)
        alpha = C.c_double()
        pt = np.zeros(2)
        s = lib().orc_segment_query_first(self._h, _p(tc), self_agent, a[0], a[1], b[0], b[1], radius,
                                          C.addressof(alpha), _p(pt))
        return s, alpha.value, pt

    def space_step(self, st: OracleState) -> None:
        """``space.step(dt)`` alone (cpSpaceStep) for every world of ``st``."""
        cs = st.c_struct()
        lib().orc_space_step(self._h, C.byref(cs))

    def point_query_nearest(self, tc: np.ndarray, self_agent: int, p, max_distance: float):
        tc = np.ascontiguousarray(tc, np.float64)
        d = C.c_double()
        s = lib().orc_point_query_nearest(self._h, _p(tc), self_agent, float(p[0]), float(p[1]), float(max_distance),
                                          C.addressof(d))
        return s, d.value

    def hull_distance(self, h: int, p) -> float:
        return lib().orc_hull_distance(self._h, h, float(p[0]), float(p[1]))


def gae(rewards, dones, values, last_values, gamma=0.99, lam=0.95, normalize=True):
    r = np.ascontiguousarray(rewards, np.float32)
    d = np.ascontiguousarray(dones, np.uint8)
    v = np.ascontiguousarray(values, np.float32)
    lv = np.ascontiguousarray(last_values, np.float32)
    T, M = r.shape
    ret = np.zeros_like(r)
    adv = np.zeros_like(r)
    lib().orc_gae(_p(r), _p(d), _p(v), _p(lv), _p(ret), _p(adv), T, M, gamma, lam, int(normalize))
    return ret, adv


def philox(c, k):
    out = np.zeros(4, np.uint32)
    lib().orc_philox(*[int(x) for x in c], *[int(x) for x in k], _p(out))
    return out


def num_threads() -> int:
    return lib().orc_num_threads()

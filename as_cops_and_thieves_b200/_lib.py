"""ctypes binding of ``libcat_b200.so`` (C ABI in ``include/cat_b200.h``).

Fails loudly: if the shared library is missing it is built with nvcc; if that is impossible, or a
compute entry point is called without a CUDA device, an exception is raised.  There is no CPU or
PyTorch fallback anywhere in this package.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

from . import build as _build

CAT_ABI_VERSION = 5
CAT_MAX_AGENTS = 8
CAT_MAX_RAYS = 128
CAT_WALL_SLOTS = 4
CAT_NEAR_SLOTS = 4


class CatError(RuntimeError):
    pass


class CatMapDesc(C.Structure):
    _fields_ = [
        ("n_hulls", C.c_int32), ("n_edges", C.c_int32),
        ("hull_off", C.c_void_p), ("vert", C.c_void_p), ("normal", C.c_void_p), ("edge_len", C.c_void_p),
        ("hull_bb", C.c_void_p),
        ("n_cops", C.c_int32), ("n_thieves", C.c_int32),
        ("init_pos", C.c_void_p), ("region_off", C.c_void_p), ("regions", C.c_void_p),
        ("grid_x0", C.c_double), ("grid_y0", C.c_double), ("cell", C.c_double),
        ("nx", C.c_int32), ("ny", C.c_int32),
        ("con_cell_off", C.c_void_p), ("con_cell_hulls", C.c_void_p),
        ("view_cell_off", C.c_void_p), ("view_cell_edges", C.c_void_p), ("view_range", C.c_double),
    ]


class CatParams(C.Structure):
    _fields_ = [
        ("dt", C.c_double), ("max_step_count", C.c_int32),
        ("unit_velocity", C.c_double), ("unit_mass", C.c_double), ("unit_size", C.c_double),
        ("max_speed", C.c_double), ("termination_radius", C.c_double),
        ("ray_length", C.c_double), ("ray_radius", C.c_double), ("wall_radius", C.c_double),
        ("n_rays", C.c_int32), ("iterations", C.c_int32),
        ("collision_slop", C.c_double), ("collision_bias", C.c_double),
        ("collision_persistence", C.c_int32), ("stale_shape_cache", C.c_int32),
        ("auto_reset", C.c_int32), ("seed", C.c_uint64), ("ray_list_cell", C.c_double),
    ]


class CatStepIO(C.Structure):
    _fields_ = [
        ("actions", C.c_void_p), ("actions_kind", C.c_int32), ("reset_mask", C.c_void_p),
        ("obs_dist", C.c_void_p), ("obs_type", C.c_void_p), ("reward", C.c_void_p),
        ("terminated", C.c_void_p), ("truncated", C.c_void_p), ("winner", C.c_void_p),
        ("shared_dist", C.c_void_p), ("shared_type", C.c_void_p), ("team_pos", C.c_void_p),
        ("obs_f32", C.c_void_p), ("state_f32", C.c_void_p), ("hit_point", C.c_void_p),
        ("obs_dist_world_stride", C.c_int32), ("obs_type_world_stride", C.c_int32),
        ("record", C.c_void_p), ("record_world_stride", C.c_int32),
        ("critic_f32", C.c_void_p), ("obs_bf16", C.c_void_p), ("critic_bf16", C.c_void_p), ("record_packed_types", C.c_int32),
    ]


class CatRecordLayout(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("bytes", "off_dist", "off_type", "off_reward", "off_terminated",
                                         "off_truncated", "off_winner", "type_bits")]


class CatEnvInfo(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "n_worlds", "n_agents", "n_cops", "n_thieves", "n_rays", "n_hulls", "n_edges", "state_dim",
        "record_words", "map_blob_bytes", "smem_bytes_per_cta", "warps_per_cta", "grid", "n_pairs",
        "ray_list_cells", "ray_list_nx", "ray_list_ny")] + [("ray_list_cell", C.c_float), ("ray_list_bytes", C.c_int64)]


class CatStateView(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "pos", "vel", "vbias", "tc", "step_count", "episode", "wall_hull", "wall_age", "wall_jn",
        "pair_age", "pair_jn")]


#: every symbol ``include/cat_b200.h`` declares
EXPORTS = (
    "cat_abi_version", "cat_last_error", "cat_env_create", "cat_env_destroy", "cat_env_info",
    "cat_env_record_layout", "cat_env_packed_record_layout", "cat_env_overflow_counts", "cat_ray_lists_host",
    "cat_env_set_seed", "cat_env_state_bytes", "cat_env_init_state", "cat_env_reset", "cat_env_step", "cat_env_step_host", "cat_env_observe",
    "cat_env_get_state", "cat_env_set_state", "cat_gae", "cat_adv_normalize",
)

_LIB = None


def lib_path() -> Path:
    return _build.LIB


def load():
    """Load (building first if needed) the CUDA extension.  Raises if it cannot be had."""
    global _LIB
    if _LIB is not None:
        return _LIB
    import os
    override = os.environ.get("CAT_B200_LIB")  # developer knob: load a tuning variant of the extension
    path = Path(override) if override else _build.build()
    L = C.CDLL(str(path))
    for name in EXPORTS:
        if not hasattr(L, name):
            raise CatError(f"{path} does not export {name}")
    L.cat_abi_version.restype = C.c_int
    if L.cat_abi_version() != CAT_ABI_VERSION:
        raise CatError(f"ABI mismatch: library {L.cat_abi_version()} != binding {CAT_ABI_VERSION}")
    L.cat_last_error.restype = C.c_char_p
    L.cat_env_create.argtypes = [C.POINTER(CatMapDesc), C.POINTER(CatParams), C.c_int32, C.c_int64, C.c_int32,
                                 C.POINTER(C.c_void_p)]
    L.cat_env_destroy.argtypes = [C.c_void_p]
    L.cat_env_info.argtypes = [C.c_void_p, C.POINTER(CatEnvInfo)]
    L.cat_env_set_seed.argtypes = [C.c_void_p, C.c_uint64]
    L.cat_env_state_bytes.restype = C.c_size_t
    L.cat_env_state_bytes.argtypes = [C.c_void_p]
    L.cat_env_init_state.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    for fn in (L.cat_env_reset, L.cat_env_step, L.cat_env_observe):
        fn.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(CatStepIO), C.c_void_p]
    L.cat_env_step_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]
    L.cat_ray_lists_host.argtypes = [C.POINTER(CatMapDesc), C.c_int32, C.c_double, C.c_double, C.c_double, C.POINTER(C.c_double * 5),
                                     C.c_void_p, C.POINTER(C.c_int64), C.c_void_p, C.POINTER(C.c_int64)]
    L.cat_env_record_layout.argtypes = [C.c_void_p, C.POINTER(CatRecordLayout)]
    L.cat_env_packed_record_layout.argtypes = [C.c_void_p, C.POINTER(CatRecordLayout)]
    L.cat_env_overflow_counts.argtypes = [C.c_void_p, C.POINTER(C.c_uint64 * 2), C.c_int32]
    for fn in (L.cat_env_get_state, L.cat_env_set_state):
        fn.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(CatStateView), C.c_void_p]
    L.cat_gae.argtypes = [C.c_void_p] * 7 + [C.c_int32, C.c_int32, C.c_float, C.c_float, C.c_void_p]
    L.cat_adv_normalize.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p]
    _LIB = L
    return L


def check(status: int, what: str) -> None:
    if status != 0:
        msg = load().cat_last_error()
        raise CatError(f"{what} failed ({status}): {msg.decode() if msg else ''}")


def np_ptr(a: np.ndarray) -> C.c_void_p:
    return a.ctypes.data_as(C.c_void_p)

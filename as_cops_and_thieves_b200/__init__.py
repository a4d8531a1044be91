"""B200-native batched cops-and-thieves environment step.

Public surface (mirrors the reference's ``src/environments`` + ``src/maps``):

* ``Map``, ``compile_map``, ``load_named_map`` — map loading / compilation (host, init-time)
* ``SimpleEnv`` / ``BaseEnv`` — single-world PettingZoo ``ParallelEnv`` face
* ``BatchedCopsThievesEnv`` — N-world skrl multi-agent-wrapper face
* ``CatWorlds`` — the thin object over the C ABI (``include/cat_b200.h``)
* ``compute_gae`` — MAPPO GAE + advantage normalisation kernels

Importing the package does not need a GPU; creating an environment does (no CPU fallback).
"""
from .maps import Map, CompiledMap, compile_map, load_named_map, free_space_regions, builtin_map_path  # noqa: F401
from .params import EnvParams, load_physical_params  # noqa: F401

__all__ = ["Map", "CompiledMap", "compile_map", "load_named_map", "free_space_regions", "builtin_map_path",
           "EnvParams", "load_physical_params", "CatWorlds", "BaseEnv", "SimpleEnv", "BatchedCopsThievesEnv",
           "compute_gae"]


def __getattr__(name):  # lazy: these import torch
    if name == "CatWorlds":
        from .worlds import CatWorlds
        return CatWorlds
    if name in ("BaseEnv", "SimpleEnv", "BatchedCopsThievesEnv"):
        from . import env
        return getattr(env, name)
    if name == "compute_gae":
        from .gae import compute_gae
        return compute_gae
    raise AttributeError(name)

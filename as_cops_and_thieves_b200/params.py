"""Physical constants of the environment.

Defaults are the reference's ``[tool.physical-params]`` table (``/root/reference/pyproject.toml:12-19``,
read through ``/root/reference/src/utils/toml_utils.py:41-46`` from the CWD on every getter call),
the sensor constants hard-coded in ``Entity.__init__`` (``/root/reference/src/agents/entity.py:84-86``),
the wall radius of ``Map.populate_space`` (``/root/reference/src/maps/map.py:127``) and the
Chipmunk2D ``cpSpace`` defaults the reference never overrides (``base_env.py:77``).
"""
from __future__ import annotations

from dataclasses import dataclass, asdict
from pathlib import Path
from typing import Optional


@dataclass
class EnvParams:
    dt: float = 1.0 / 60.0                 # simple_env.py:20
    max_step_count: int = 400              # simple_env.py:19
    unit_velocity: float = 10.0            # pyproject.toml:13
    unit_mass: float = 1.0                 # :14
    unit_size: float = 5.0                 # :15
    max_speed: float = 125.0               # :16
    termination_radius: float = 20.0       # :19
    ray_length: float = 400.0              # entity.py:84
    ray_radius: float = 1.0                # entity.py:196 (segment_query_first radius)
    wall_radius: float = 1.0               # map.py:127
    n_rays: int = 90                       # entity.py:86
    iterations: int = 10                   # cpSpace default
    collision_slop: float = 0.1            # cpSpace default
    collision_bias: float = (1.0 - 0.1) ** 60.0   # cpSpace default
    collision_persistence: int = 3         # cpSpace default
    stale_shape_cache: int = 1             # pymunk behaviour, SURVEY.md A.10 / C-4
    auto_reset: int = 1                    # batched env: re-spawn inside the step
    seed: int = 0
    ray_list_cell: float = 0.0             # sensor-sweep candidate lists: cell size (0 = automatic, < 0 = rasterise instead)

    def as_dict(self) -> dict:
        return asdict(self)


def load_physical_params(pyproject: Optional[str] = None) -> dict:
    """Read ``[tool.physical-params]`` from a pyproject.toml like the reference's getters do.

    Returns only the keys the environment uses; missing file -> the reference's shipped values.
    """
    out = dict(unit_velocity=10.0, unit_mass=1.0, unit_size=5.0, max_speed=125.0, termination_radius=20.0)
    path = Path(pyproject) if pyproject else Path("pyproject.toml")
    if path.exists():
        try:
            import tomllib as _toml
        except ImportError:  # pragma: no cover
            import tomli as _toml
        with open(path, "rb") as f:
            cfg = _toml.load(f)
        table = cfg.get("tool", {}).get("physical-params", {})
        for k in out:
            if k in table:
                out[k] = float(table[k])
    return out

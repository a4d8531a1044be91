"""Observation/action space objects.

Uses gymnasium when it is installed (the reference does, ``entity.py:88-107``); otherwise a
minimal stand-in with the attributes skrl's wrappers and the reference's model builders read
(``shape``, ``dtype``, ``low``/``high``, ``n``, ``spaces``, key order).  ``Dict`` sorts its keys
like gymnasium's does — that ordering fixes the flattened observation layout.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Mapping

import numpy as np

try:  # pragma: no cover - exercised only where gymnasium exists
    from gymnasium.spaces import Box, Dict, Discrete  # type: ignore
    HAVE_GYMNASIUM = True
except Exception:  # ModuleNotFoundError in this image
    HAVE_GYMNASIUM = False

    class Space:
        shape = None
        dtype = None

        def contains(self, x) -> bool:  # pragma: no cover - trivial
            raise NotImplementedError

        def __contains__(self, x) -> bool:
            return self.contains(x)

    class Box(Space):
        def __init__(self, low, high, shape=None, dtype=np.float32):
            self.dtype = np.dtype(dtype)
            if shape is None:
                shape = np.broadcast(np.asarray(low), np.asarray(high)).shape
            self.shape = tuple(int(s) for s in shape)
            self.low = np.full(self.shape, low, dtype=self.dtype) if np.isscalar(low) else np.asarray(low, self.dtype)
            self.high = np.full(self.shape, high, dtype=self.dtype) if np.isscalar(high) else np.asarray(high, self.dtype)

        def contains(self, x) -> bool:
            x = np.asarray(x)
            return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

        def sample(self):
            return np.random.uniform(self.low.astype(np.float64), self.high.astype(np.float64)).astype(self.dtype)

        def __repr__(self):
            return f"Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})"

        def __eq__(self, other):
            return isinstance(other, Box) and self.shape == other.shape and self.dtype == other.dtype \
                and np.array_equal(self.low, other.low) and np.array_equal(self.high, other.high)

    class Discrete(Space):
        def __init__(self, n: int, start: int = 0):
            self.n, self.start = int(n), int(start)
            self.shape, self.dtype = (), np.dtype(np.int64)

        def contains(self, x) -> bool:
            return self.start <= int(x) < self.start + self.n

        def sample(self):
            return int(np.random.randint(self.start, self.start + self.n))

        def __repr__(self):
            return f"Discrete({self.n})"

        def __eq__(self, other):
            return isinstance(other, Discrete) and self.n == other.n and self.start == other.start

    class Dict(Space):
        def __init__(self, spaces: Mapping):
            try:
                items = sorted(spaces.items())  # gymnasium keeps plain-dict keys sorted
            except TypeError:
                items = list(spaces.items())
            self.spaces = OrderedDict(items)

        def __getitem__(self, k):
            return self.spaces[k]

        def __iter__(self):
            return iter(self.spaces)

        def __len__(self):
            return len(self.spaces)

        def keys(self):
            return self.spaces.keys()

        def values(self):
            return self.spaces.values()

        def items(self):
            return self.spaces.items()

        def contains(self, x) -> bool:
            return isinstance(x, Mapping) and set(x.keys()) == set(self.spaces.keys()) and \
                all(self.spaces[k].contains(v) for k, v in x.items())

        def sample(self):
            return {k: s.sample() for k, s in self.spaces.items()}

        def __repr__(self):
            return "Dict(" + ", ".join(f"{k!r}: {s!r}" for k, s in self.spaces.items()) + ")"

        def __eq__(self, other):
            return isinstance(other, Dict) and list(self.spaces.items()) == list(other.spaces.items())


def flatdim(space) -> int:
    """Number of float32 entries skrl's ``flatten_tensorized_space`` yields for ``space``."""
    if isinstance(space, Dict):
        return sum(flatdim(s) for s in space.spaces.values())
    if isinstance(space, Discrete):
        return 1
    return int(np.prod(space.shape))


def flatten(space, x) -> np.ndarray:
    """Flatten one sample of ``space`` to float32 in skrl / gymnasium key order."""
    if isinstance(space, Dict):
        return np.concatenate([flatten(s, x[k]) for k, s in space.spaces.items()])
    return np.asarray(x, dtype=np.float32).reshape(-1)

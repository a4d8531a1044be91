"""Host-side ``rgb_array`` renderer (replacement for ``base_env.py:481-510``).

Visual only — never on the hot path.  Returns an array shaped like
``pygame.surfarray.array3d(window)``: ``(width, height, 3)`` uint8, walls white on black
(``pyproject.toml:21-23``), cops blue, thieves red (``cop.py:30``, ``thief.py:30``).
"""
from __future__ import annotations

import numpy as np


def render_rgb(cmap, pos: np.ndarray, agent_radius: float = 5.0) -> np.ndarray:
    W, H = int(cmap.window[0]), int(cmap.window[1])
    img = np.zeros((W, H, 3), np.uint8)
    for h in range(cmap.n_hulls):
        o, e = cmap.hull_off[h], cmap.hull_off[h + 1]
        l, b, r, t = cmap.hull_bb[h]
        x0, x1 = max(0, int(np.floor(l - 1))), min(W, int(np.ceil(r + 1)) + 1)
        y0, y1 = max(0, int(np.floor(b - 1))), min(H, int(np.ceil(t + 1)) + 1)
        if x1 <= x0 or y1 <= y0:
            continue
        xs, ys = np.meshgrid(np.arange(x0, x1) + 0.5, np.arange(y0, y1) + 0.5, indexing="ij")
        inside = np.ones(xs.shape, bool)
        for (vx, vy), (nx, ny) in zip(cmap.vert[o:e], cmap.normal[o:e]):
            inside &= ((xs - vx) * nx + (ys - vy) * ny) <= 1.0  # wall radius 1
        img[x0:x1, y0:y1][inside] = (255, 255, 255)
    for a, (px, py) in enumerate(np.asarray(pos, np.float64)):
        color = (0, 0, 255) if a < cmap.n_cops else (255, 0, 0)
        x0, x1 = max(0, int(px - agent_radius - 1)), min(W, int(px + agent_radius + 2))
        y0, y1 = max(0, int(py - agent_radius - 1)), min(H, int(py + agent_radius + 2))
        if x1 <= x0 or y1 <= y0:
            continue
        xs, ys = np.meshgrid(np.arange(x0, x1) + 0.5, np.arange(y0, y1) + 0.5, indexing="ij")
        img[x0:x1, y0:y1][(xs - px) ** 2 + (ys - py) ** 2 <= agent_radius ** 2] = color
    return img

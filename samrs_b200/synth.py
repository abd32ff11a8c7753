"""Seeded synthetic remote-sensing tiles and prompts.

One definition shared by the golden generator, the parity tests and bench.py so
that every arm sees byte-identical inputs.
"""
from __future__ import annotations

import numpy as np


def tile(idx: int, size: int = 1024) -> np.ndarray:
    """uint8 HWC tile; content is irrelevant to the FLOP count, only the seed matters."""
    return np.random.default_rng(idx).integers(0, 256, (size, size, 3), dtype=np.uint8)


def hboxes(idx: int, n: int = 32, size: int = 1024, tiny: bool = False) -> np.ndarray:
    """float32 (n,4) xyxy horizontal boxes. `tiny` = the SOTA-density small-object mix of config 4."""
    rng = np.random.default_rng(1_000_003 * (idx + 1))
    c = rng.uniform(64, size - 64, (n, 2))
    wh = rng.uniform(4, 32, (n, 2)) if tiny else rng.uniform(8, 200, (n, 2))
    b = np.concatenate([c - wh / 2, c + wh / 2], axis=1)
    return np.clip(b, 0, size - 1).astype(np.float32)


def labels(idx: int, n: int = 32, classes: int = 18) -> np.ndarray:
    """int64 class ids in [0, classes): DOTA2_0 has 18 (Generate Dataset/mapping.py:46-50)."""
    return np.random.default_rng(7_000_003 * (idx + 1)).integers(0, classes, n)


def rbox_polys(idx: int, n: int = 32, size: int = 1024) -> np.ndarray:
    """(n,4,2) float32 rotated-box polygons (vertex order of `obb2poly_np_le90`,
    Generate Dataset/utils/transform.py:193-216: c-v1-v2, c+v1-v2, c+v1+v2, c-v1+v2), clipped to the tile."""
    rng = np.random.default_rng(3_000_017 * (idx + 1))
    c = rng.uniform(128, size - 128, (n, 2))
    w = rng.uniform(16, 200, n)
    h = rng.uniform(16, 200, n)
    th = rng.uniform(-np.pi / 2, np.pi / 2, n)
    cs, sn = np.cos(th), np.sin(th)
    v1 = np.stack([w / 2 * cs, w / 2 * sn], 1)
    v2 = np.stack([-h / 2 * sn, h / 2 * cs], 1)
    pts = np.stack([c - v1 - v2, c + v1 - v2, c + v1 + v2, c - v1 + v2], axis=1)
    return np.clip(pts, 0, size - 1).astype(np.float32)


def rboxes_5pt(idx: int, n: int = 32, size: int = 1024) -> np.ndarray:
    """(n,5,2) float32 point prompts: 4 rotated-box vertices + centre (BASELINE.json config 3;
    vertex math as `obb2poly_np_le90`, Generate Dataset/utils/transform.py:193-216)."""
    rng = np.random.default_rng(3_000_017 * (idx + 1))
    c = rng.uniform(128, size - 128, (n, 2))
    w = rng.uniform(16, 200, n)
    h = rng.uniform(16, 200, n)
    th = rng.uniform(-np.pi / 2, np.pi / 2, n)
    cs, sn = np.cos(th), np.sin(th)
    vx = np.stack([w / 2 * cs, w / 2 * sn], 1)
    vy = np.stack([-h / 2 * sn, h / 2 * cs], 1)
    pts = np.stack([c + vx + vy, c + vx - vy, c - vx - vy, c - vx + vy, c], axis=1)
    return np.clip(pts, 0, size - 1).astype(np.float32)


def mask_prompts(idx: int, n: int = 4) -> np.ndarray:
    """(n,1,256,256) float32 +-1000 mask prompts, the modality of
    `Generate Dataset/main_sam_rbox_mask_instance.py:125-141,159-164`
    (axis-aligned rectangles here; rasterisation itself stays on the host)."""
    rng = np.random.default_rng(5_000_011 * (idx + 1))
    m = np.full((n, 1, 256, 256), -1000.0, dtype=np.float32)
    for j in range(n):
        x0, y0 = rng.integers(8, 160, 2)
        w, h = rng.integers(8, 90, 2)
        m[j, 0, y0:y0 + h, x0:x0 + w] = 1000.0
    return m

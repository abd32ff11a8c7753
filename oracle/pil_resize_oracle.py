"""ORACLE (test infrastructure, not a product path): numpy restatement of Pillow's 8-bit bilinear resize.

The reference resizes the image with `torchvision.transforms.functional.resize(to_pil_image(image), target_size)`
(`segment_anything/utils/transforms.py:30-31`), i.e. `PIL.Image.resize(..., BILINEAR)`.  Pillow (an un-vendored
dependency; installed here: 12.2.0; the reference pins none) implements it in `src/libImaging/Resample.c` as two
separable passes over uint8 data with 22-bit fixed-point coefficients:

  * per output index: centre = (i + 0.5) * scale, support = max(scale, 1) (the triangle filter widens when shrinking),
    taps [xmin, xmax) = [int(centre - support + 0.5), int(centre + support + 0.5)) clipped to the input,
    weights triangle((x + xmin - centre + 0.5) / max(scale, 1)) normalised to sum 1 (all in double);
  * weights -> int via int(+-0.5 + w * 2^22); every pass accumulates 2^21 + sum(pixel * weight) in int32, shifts right by
    22 and clamps to [0, 255]; the horizontal pass runs first and its uint8 result feeds the vertical pass.

Pinned against Pillow itself in `tests/test_resize.py` (bit-exact on every tested size, up- and down-scaling).
"""
from __future__ import annotations

import math
from typing import Tuple

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def coeffs(in_size: int, out_size: int) -> Tuple[np.ndarray, np.ndarray]:
    """(bounds int32 [out, 2] = (first tap, tap count), weights int32 [out, ksize])."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = np.zeros(ksize, np.float64)
        for x in range(xmax):
            v = abs((x + xmin - center + 0.5) * ss)
            w[x] = 1.0 - v if v < 1.0 else 0.0
        # Pillow sums sequentially in double; so does this loop (np.sum pairwise could differ in the last bit)
        ww = 0.0
        for x in range(xmax):
            ww += w[x]
        if ww != 0.0:
            w[:xmax] /= ww
        for x in range(ksize):
            kk[xx, x] = int(-0.5 + w[x] * (1 << PRECISION_BITS)) if w[x] < 0 else int(0.5 + w[x] * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def _pass(img: np.ndarray, bounds: np.ndarray, kk: np.ndarray) -> np.ndarray:
    """Resample axis 0 of a (n, ...) uint8 array."""
    out = np.empty((bounds.shape[0],) + img.shape[1:], np.uint8)
    src = img.astype(np.int64)
    for i, (lo, n) in enumerate(bounds):
        acc = np.full(img.shape[1:], 1 << (PRECISION_BITS - 1), np.int64)
        for k in range(n):
            acc += src[lo + k] * int(kk[i, k])
        out[i] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return out


def resize_bilinear_u8(img: np.ndarray, out_hw: Tuple[int, int]) -> np.ndarray:
    """(H, W, C) uint8 -> (out_h, out_w, C) uint8, as `Image.fromarray(img).resize((out_w, out_h), Image.BILINEAR)`."""
    h, w = img.shape[:2]
    oh, ow = out_hw
    cur = img
    if ow != w:                                                   # horizontal first (Resample.c ImagingResampleInner)
        bh, kh = coeffs(w, ow)
        cur = np.swapaxes(_pass(np.swapaxes(cur, 0, 1), bh, kh), 0, 1)
    if oh != h:
        bv, kv = coeffs(h, oh)
        cur = _pass(cur, bv, kv)
    return np.ascontiguousarray(cur)

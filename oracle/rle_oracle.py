"""ORACLE (test infrastructure, not a product path): numpy restatement of the reference's run-length semantics.

Follows `segment_anything/utils/amg.py:107-135` (`mask_to_rle_pytorch`): masks are flattened in Fortran order
(`permute(0, 2, 1).flatten(1)`, :114-115), change indices are where neighbouring pixels differ (:118-119), the runs are
the differences of `[0, change+1..., h*w]` (:125-132) and a mask whose first pixel is set gets a leading 0 (:133).
Area is the driver's `np.sum(mask)` (`Generate Dataset/main_sam_hbox_semantic.py:203`).

Pinned against the reference function itself: `oracle/make_golden_rle.py` imports `mask_to_rle_pytorch` from
/root/reference and stores its output for the masks in `tests/golden/rle_cases.npz`.
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np


def mask_to_rle(mask: np.ndarray) -> List[int]:
    """Uncompressed COCO counts of one (h, w) boolean mask."""
    h, w = mask.shape
    flat = np.asarray(mask, dtype=bool).T.reshape(-1)             # amg.py:114-115 (Fortran order)
    change = np.nonzero(flat[1:] ^ flat[:-1])[0]                   # amg.py:118-119
    idx = np.concatenate([[0], change + 1, [h * w]])               # amg.py:125-131
    counts = [] if not flat[0] else [0]                            # amg.py:133
    counts.extend((idx[1:] - idx[:-1]).tolist())                   # amg.py:132,134
    return counts


def encode_batch(masks: np.ndarray) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """(counts concatenated int64, offsets int64 [B+1], area int64 [B]) of (B, h, w) masks."""
    counts, offsets, area = [], [0], []
    for m in masks:
        c = mask_to_rle(m)
        counts.extend(c)
        offsets.append(len(counts))
        area.append(int(np.sum(m)))                                # main_sam_hbox_semantic.py:203
    return np.asarray(counts, dtype=np.int64), np.asarray(offsets, dtype=np.int64), np.asarray(area, dtype=np.int64)

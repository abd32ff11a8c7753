"""Output writers of the generation drivers: the on-disk contract the training half reads.

`Generate Dataset/main_sam_hbox_semantic.py:211-216` saves, per image,
  gray/<name>.png   uint8 class id per pixel, 255 = ignore (`seg_mask`, :162,197)
  color/<name>.png  RGB rendering through the dataset's `MAPPING` colour table (`seg_color`, :163,198; `mapping.py:3-42`)
  ins/<name>.pkl    pickle of a list of `{mask (COCO RLE), bbox, category, label, size}` dicts (:200-205)
This module writes the same three files from what the engine produces on the device: the fused label map
(`Engine.semantic_reduce`) and the instance payload (`Engine.rle_encode` -> `samrs_b200.rle.instance_records`).
The colour table is passed in (the drivers import it from their own `mapping.py`), so no dataset constants live here.
"""
from __future__ import annotations

import os
import pickle
from typing import Any, Dict, List, Mapping, Sequence, Tuple

import numpy as np


def color_lut(mapping: Mapping[int, Sequence[int]]) -> np.ndarray:
    """(256, 3) uint8 lookup table from a `{class id: (r, g, b)}` dict; ids absent from the dict render black, which is
    what the driver's zero-initialised `seg_color` shows for them (it is only ever written through MAPPING)."""
    lut = np.zeros((256, 3), dtype=np.uint8)
    for k, rgb in mapping.items():
        if not 0 <= int(k) <= 255:
            raise ValueError(f"class id {k} does not fit the uint8 label map")
        lut[int(k)] = np.asarray(rgb, dtype=np.uint8)
    return lut


def colorize(label_map: np.ndarray, mapping: Mapping[int, Sequence[int]], background: Sequence[int] = (255, 255, 255)) -> np.ndarray:
    """`seg_color` of the drivers: 255-initialised canvas (`main_sam_hbox_semantic.py:163`), painted pixels take
    MAPPING[label] (:198).  Unpainted pixels are the ones whose label is still 255."""
    if label_map.dtype != np.uint8 or label_map.ndim != 2:
        raise ValueError("label map must be a 2-D uint8 array")
    lut = color_lut(mapping)
    lut[255] = np.asarray(background, dtype=np.uint8)
    return lut[label_map]


def save_tile(save_dir: str, name: str, label_map: np.ndarray, mapping: Mapping[int, Sequence[int]],
              instances: List[Dict[str, Any]]) -> Tuple[str, str, str]:
    """Writes gray/<name>.png, color/<name>.png and ins/<name>.pkl under `save_dir` (directories are created)."""
    from PIL import Image
    paths = []
    for sub in ("gray", "color", "ins"):
        os.makedirs(os.path.join(save_dir, sub), exist_ok=True)
    p = os.path.join(save_dir, "gray", name + ".png")
    Image.fromarray(np.ascontiguousarray(label_map)).save(p)
    paths.append(p)
    p = os.path.join(save_dir, "color", name + ".png")
    Image.fromarray(colorize(label_map, mapping)).save(p)
    paths.append(p)
    p = os.path.join(save_dir, "ins", name + ".pkl")
    with open(p, "wb") as f:
        pickle.dump(instances, f)
    paths.append(p)
    return tuple(paths)

"""CPU oracle for the rotated-box -> mask-prompt rasterisation.  TEST INFRASTRUCTURE ONLY.

Restates `Generate Dataset/main_sam_rbox_mask_instance.py:125-141` with the same OpenCV calls the driver makes
(the arithmetic lives in the un-vendored dependency OpenCV; installed here: 4.13.0, the reference pins none):

    canvas  = zeros(H, W, 3) uint8;  cv2.fillPoly(canvas, [poly.astype(int32)], (255,255,255))         :127-128
    box_mask = -1000 everywhere, +1000 where all three channels are 255 (float64)                       :129-132
    box_mask = cv2.resize(box_mask, (new_w, new_h), INTER_LINEAR)   long side -> 1024                   :134-135
    box_mask = cv2.copyMakeBorder(..., bottom = 1024 - new_h, right = 1024 - new_w, value = -1000)      :136-138
    box_mask = cv2.resize(box_mask, (256, 256), INTER_LINEAR)                                           :139
    torch.tensor(box_mask).float()                                                                      :140

Parity pin: the functions below ARE the reference recipe executed by cv2 itself; `samrs_b200`'s device rasteriser is
compared with them (tests/test_rbox_prompt.py) and fixtures of their output are committed by oracle/make_golden_h.py.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np


def preprocess_shape(oldh: int, oldw: int, long_side: int = 1024) -> Tuple[int, int]:
    """`ResizeLongestSide.get_preprocess_shape` (SA/utils/transforms.py:94-102)."""
    scale = long_side * 1.0 / max(oldh, oldw)
    return int(oldh * scale + 0.5), int(oldw * scale + 0.5)


def mask_prompt(poly: np.ndarray, image_hw: Tuple[int, int], img_size: int = 1024) -> np.ndarray:
    """One (4,2) polygon in original-image pixels -> (256,256) float32 mask prompt (+-1000 with blended edges)."""
    import cv2
    H, W = int(image_hw[0]), int(image_hw[1])
    canvas = np.zeros([H, W, 3], dtype=np.uint8)
    box_pts = np.asarray(poly).astype(np.int32)
    draws = cv2.fillPoly(canvas, [box_pts], color=(255, 255, 255))
    box_mask = -1000 * np.ones([H, W])
    posi = np.all(draws == np.array([255, 255, 255]).reshape(1, 1, 3), axis=2)
    box_mask[posi] = 1000
    target = preprocess_shape(H, W, img_size)
    box_mask = cv2.resize(box_mask, target[::-1], interpolation=cv2.INTER_LINEAR)
    padh, padw = img_size - box_mask.shape[0], img_size - box_mask.shape[1]
    box_mask = cv2.copyMakeBorder(box_mask, 0, padh, 0, padw, cv2.BORDER_CONSTANT, value=-1000)
    box_mask = cv2.resize(box_mask, (256, 256), interpolation=cv2.INTER_LINEAR)
    return box_mask.astype(np.float32)          # == torch.tensor(box_mask).float()


def mask_prompts(polys: np.ndarray, image_hw: Tuple[int, int], img_size: int = 1024) -> np.ndarray:
    """(n,4,2) polygons -> (n,1,256,256) float32, the `mask_input` the driver passes to predict_torch (:159-164)."""
    return np.stack([mask_prompt(p, image_hw, img_size) for p in polys], 0)[:, None]


def fill_mask(poly: np.ndarray, image_hw: Tuple[int, int]) -> np.ndarray:
    """The boolean `posi` raster of cv2.fillPoly alone (stage-wise check of the device rasteriser)."""
    import cv2
    canvas = np.zeros([int(image_hw[0]), int(image_hw[1]), 3], dtype=np.uint8)
    draws = cv2.fillPoly(canvas, [np.asarray(poly).astype(np.int32)], color=(255, 255, 255))
    return np.all(draws == 255, axis=2)

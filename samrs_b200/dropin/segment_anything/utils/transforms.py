"""Prompt / image geometry of the predictor: scale everything so that the long image side becomes `target_length`.

Same public surface and results as the reference class (segment_anything/utils/transforms.py:16-102), which the drivers
reach through `predictor.transform.apply_boxes_torch` (main_sam_hbox_semantic.py:174).  All of it is cheap host-side
arithmetic and the identity on 1024x1024 tiles; the image itself is resized with PIL's bilinear filter, which is what
torchvision's `resize(to_pil_image(...))` does in the reference (:30-31).
"""
from typing import Tuple

import numpy as np
import torch
from PIL import Image


def _target_hw(h: int, w: int, long_side: int) -> Tuple[int, int]:
    """Rounded size after scaling the longer side to `long_side` (reference :94-102: half-up rounding via int(x + 0.5))."""
    s = long_side * 1.0 / max(h, w)
    return int(h * s + 0.5), int(w * s + 0.5)


class ResizeLongestSide:
    def __init__(self, target_length: int) -> None:
        self.target_length = target_length

    # ------------------------------------------------------------------ helpers
    def _factors(self, original_size: Tuple[int, ...]) -> Tuple[float, float]:
        """(x factor, y factor) as Python floats, formed exactly like the reference's `new_w / old_w`, `new_h / old_h`."""
        h0, w0 = original_size
        h1, w1 = _target_hw(h0, w0, self.target_length)
        return w1 / w0, h1 / h0

    @staticmethod
    def get_preprocess_shape(oldh: int, oldw: int, long_side_length: int) -> Tuple[int, int]:
        return _target_hw(oldh, oldw, long_side_length)

    # ------------------------------------------------------------------ image
    def apply_image(self, image: np.ndarray) -> np.ndarray:
        h1, w1 = _target_hw(image.shape[0], image.shape[1], self.target_length)
        if image.shape[0] == h1 and image.shape[1] == w1:
            return np.ascontiguousarray(image)                       # nothing to resample (every 1024x1024 tile)
        return np.array(Image.fromarray(image).resize((w1, h1), Image.BILINEAR))

    # ------------------------------------------------------------------ numpy prompts
    def apply_coords(self, coords: np.ndarray, original_size: Tuple[int, ...]) -> np.ndarray:
        fx, fy = self._factors(original_size)
        out = np.array(coords, dtype=float, copy=True)
        out[..., 0] *= fx
        out[..., 1] *= fy
        return out

    def apply_boxes(self, boxes: np.ndarray, original_size: Tuple[int, ...]) -> np.ndarray:
        corners = self.apply_coords(boxes.reshape(-1, 2, 2), original_size)
        return corners.reshape(-1, 4)

    # ------------------------------------------------------------------ torch prompts (stay on their device)
    def apply_coords_torch(self, coords: torch.Tensor, original_size: Tuple[int, ...]) -> torch.Tensor:
        fx, fy = self._factors(original_size)
        out = coords.detach().clone().to(torch.float)
        out[..., 0].mul_(fx)
        out[..., 1].mul_(fy)
        return out

    def apply_boxes_torch(self, boxes: torch.Tensor, original_size: Tuple[int, ...]) -> torch.Tensor:
        corners = self.apply_coords_torch(boxes.reshape(-1, 2, 2), original_size)
        return corners.reshape(-1, 4)

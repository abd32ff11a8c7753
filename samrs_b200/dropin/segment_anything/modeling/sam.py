"""`Sam` handle with the attribute surface the SAMRS drivers and `SamPredictor` touch
(reference: segment_anything/modeling/sam.py:18-50, build_sam.py:55-107)."""
from __future__ import annotations

from types import SimpleNamespace
from typing import Any, Dict, Optional

import torch

from samrs_b200.config import SamGeometry
from samrs_b200.weights import check_state_dict


class Sam:
    mask_threshold: float = 0.0
    image_format: str = "RGB"

    def __init__(self, geometry: SamGeometry):
        self.geometry = geometry
        # drivers read sam.image_encoder.img_size (main_sam_rbox_mask_instance.py:135-138)
        self.image_encoder = SimpleNamespace(img_size=geometry.img_size)
        self._state: Optional[Dict[str, torch.Tensor]] = None
        self._device = torch.device("cpu")
        self.engine = None
        self.training = False

    # -- nn.Module-like surface ---------------------------------------------------
    @property
    def device(self) -> Any:
        return self._device

    def eval(self) -> "Sam":
        return self

    def load_state_dict(self, state_dict: Dict[str, torch.Tensor], strict: bool = True) -> None:
        check_state_dict(self.geometry, state_dict)
        self._state = state_dict
        if self.engine is not None:
            self.engine.load_state_dict(state_dict)

    def state_dict(self) -> Dict[str, torch.Tensor]:
        if self._state is None:
            raise RuntimeError("no weights loaded")
        return self._state

    def to(self, device=None, **kwargs) -> "Sam":
        device = torch.device(device if device is not None else kwargs.get("device", "cuda"))
        if device.type != "cuda":
            raise RuntimeError("samrs_b200 has no CPU path: move the model to a CUDA device")
        from samrs_b200.engine import Engine
        if device.index is None:                                   # "cuda" means the current device, as in torch
            device = torch.device("cuda", torch.cuda.current_device())
        if self.engine is None or self.engine.device != device:
            self.engine = Engine(self.geometry, device)
            if self._state is not None:
                self.engine.load_state_dict(self._state)
        self._device = self.engine.device
        return self

    def cuda(self, device=None) -> "Sam":
        return self.to(torch.device("cuda", device) if isinstance(device, int) else (device or "cuda"))

    def _require_engine(self):
        if self.engine is None:
            raise RuntimeError("Sam must be moved to a CUDA device with .to(device=...) before use (no CPU path)")
        if not self.engine.weights_loaded:
            raise RuntimeError("Sam has no weights: pass checkpoint= to the registry builder or call load_state_dict")
        return self.engine

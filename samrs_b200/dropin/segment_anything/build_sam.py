"""`sam_model_registry` and builders (reference: segment_anything/build_sam.py:14-52,102-107)."""
import torch

from samrs_b200.config import geometry

from .modeling import Sam


def _build_sam(variant: str, checkpoint=None) -> Sam:
    sam = Sam(geometry(variant))
    sam.eval()
    if checkpoint is not None:
        with open(checkpoint, "rb") as f:
            state_dict = torch.load(f, map_location="cpu")
        sam.load_state_dict(state_dict)
    return sam


def build_sam_vit_h(checkpoint=None):
    return _build_sam("vit_h", checkpoint)


build_sam = build_sam_vit_h


def build_sam_vit_l(checkpoint=None):
    return _build_sam("vit_l", checkpoint)


def build_sam_vit_b(checkpoint=None):
    return _build_sam("vit_b", checkpoint)


sam_model_registry = {
    "default": build_sam_vit_h,
    "vit_h": build_sam_vit_h,
    "vit_l": build_sam_vit_l,
    "vit_b": build_sam_vit_b,
}

"""Drop-in replacement for the vendored `segment_anything` package of SAMRS
(`/root/reference/Generate Dataset/segment_anything/__init__.py:7-15`): the same names, backed by the
samrs_b200 CUDA engine instead of nn.Modules.  Put this directory's parent first on `sys.path` and the
drivers (`main_sam_hbox_semantic.py` & co.) import it unchanged."""
from .build_sam import (
    build_sam,
    build_sam_vit_h,
    build_sam_vit_l,
    build_sam_vit_b,
    sam_model_registry,
)
from .predictor import SamPredictor
from .automatic_mask_generator import SamAutomaticMaskGenerator

__all__ = ["build_sam", "build_sam_vit_h", "build_sam_vit_l", "build_sam_vit_b", "sam_model_registry",
           "SamPredictor", "SamAutomaticMaskGenerator"]

"""No-GPU checks of the C-ABI boundary: the in-tree library loads, exports every symbol that
include/samrs_b200.h declares, the ctypes table covers them, and the product path fails loudly without a GPU."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "samrs_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(samrs_\w+)\s*\(", text)))


def test_header_declares_the_documented_entry_points():
    syms = declared_symbols()
    for must in ("samrs_create", "samrs_load_weights", "samrs_encode", "samrs_decode", "samrs_postprocess",
                 "samrs_semantic_reduce", "samrs_set_features", "samrs_last_error", "samrs_destroy"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    from samrs_b200.engine import ABI, load_library
    lib = load_library()
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} declared in include/samrs_b200.h but not exported"
        assert name in ABI, f"{name} has no ctypes signature in samrs_b200.engine.ABI"
    for name in ABI:
        assert name in declared_symbols(), f"{name} bound in ABI but not declared in the header"


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    from samrs_b200.engine import Engine, load_library
    lib = load_library()
    h = ctypes.c_void_p()
    gi = (ctypes.c_int * 1)(1)
    rc = lib.samrs_create(0, 128, 2, 2, gi, 1, ctypes.byref(h))
    assert rc != 0 and not h.value
    assert b"no CUDA device" in lib.samrs_last_error(None)
    with pytest.raises(RuntimeError):
        Engine("vit_t64", "cpu")
    with pytest.raises(Exception):
        Engine("vit_t64", "cuda:0")


def test_dropin_package_surface_without_gpu():
    """Names and error behaviour of the drop-in `segment_anything` (reference: segment_anything/__init__.py:7-15,
    predictor.py:213-214, build_sam.py:47-52)."""
    import sys

    import samrs_b200
    sys.path.insert(0, samrs_b200.DROPIN_PATH)
    try:
        for k in [k for k in sys.modules if k == "segment_anything" or k.startswith("segment_anything.")]:
            del sys.modules[k]
        import segment_anything as sa
        assert set(sa.sam_model_registry) == {"default", "vit_h", "vit_l", "vit_b"}
        for n in ("build_sam", "build_sam_vit_h", "build_sam_vit_l", "build_sam_vit_b", "SamPredictor", "SamAutomaticMaskGenerator"):
            assert hasattr(sa, n)
        sam = sa.sam_model_registry["vit_b"]()
        assert sam.image_encoder.img_size == 1024 and sam.mask_threshold == 0.0 and sam.image_format == "RGB"
        with pytest.raises(RuntimeError):
            sam.to(device="cpu")
        pred = sa.SamPredictor(sam)
        with pytest.raises(RuntimeError, match="An image must be set"):
            pred.predict_torch(None, None, boxes=torch.zeros(1, 4))
        with pytest.raises(RuntimeError, match="An image must be set"):
            pred.get_image_embedding()
        with pytest.raises(NotImplementedError):
            sa.SamAutomaticMaskGenerator(sam)
        with pytest.raises(RuntimeError):       # strict load: wrong keys
            sam.load_state_dict({"bogus": torch.zeros(1)})
    finally:
        sys.path.remove(samrs_b200.DROPIN_PATH)
        for k in [k for k in sys.modules if k == "segment_anything" or k.startswith("segment_anything.")]:
            del sys.modules[k]

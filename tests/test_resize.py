"""Pillow-exact image resize.  CPU: the oracle restatement against Pillow itself (the reference's
`ResizeLongestSide.apply_image` is `PIL.Image.resize(BILINEAR)` through torchvision).  GPU: `samrs_resize_bilinear_u8`
against the oracle, and `SamPredictor.set_image` on a non-1024 image against encoding the PIL-resized image."""
import sys

import numpy as np
import pytest
import torch
from PIL import Image

from oracle.pil_resize_oracle import coeffs, resize_bilinear_u8

SIZES = [((800, 800), (1024, 1024)), ((613, 977), (643, 1024)), ((2000, 1500), (1024, 768)), ((300, 200), (1024, 683)),
         ((37, 500), (76, 1024)), ((1500, 1024), (1024, 699)), ((64, 64), (20, 33)), ((1024, 1000), (1024, 1000))]


def _img(h, w, seed=0):
    return np.random.default_rng(seed).integers(0, 256, (h, w, 3), dtype=np.uint8)


@pytest.mark.parametrize("src,dst", SIZES)
def test_oracle_equals_pillow(src, dst):
    img = _img(*src)
    want = np.array(Image.fromarray(img).resize((dst[1], dst[0]), Image.BILINEAR))
    assert np.array_equal(resize_bilinear_u8(img, dst), want)


def test_coefficient_tables_are_normalised():
    for n_in, n_out in [(800, 1024), (2000, 1024), (64, 20)]:
        bounds, kk = coeffs(n_in, n_out)
        assert bounds[:, 0].min() >= 0 and (bounds[:, 0] + bounds[:, 1]).max() <= n_in
        assert np.abs(kk.sum(1) - (1 << 22)).max() <= kk.shape[1]            # weights sum to one up to rounding


@pytest.fixture(scope="module")
def eng():
    from samrs_b200.engine import Engine
    return Engine("vit_t64", "cuda:0")


@pytest.mark.gpu
@pytest.mark.parametrize("src,dst", SIZES)
def test_device_resize_bit_exact(eng, src, dst):
    img = _img(*src, seed=3)
    got = eng.resize_image(torch.from_numpy(img).cuda(), dst)
    torch.cuda.synchronize()
    assert np.array_equal(got.cpu().numpy(), resize_bilinear_u8(img, dst))


@pytest.mark.gpu
def test_set_image_resizes_on_device_like_the_reference():
    import samrs_b200
    from samrs_b200.weights import synthetic_state_dict
    sys.path.insert(0, samrs_b200.DROPIN_PATH)
    try:
        for k in [k for k in sys.modules if k == "segment_anything" or k.startswith("segment_anything.")]:
            del sys.modules[k]
        from segment_anything import SamPredictor
        from segment_anything.modeling import Sam
        from samrs_b200.config import geometry
        sam = Sam(geometry("vit_t64"))
        sam.load_state_dict(synthetic_state_dict("vit_t64", 0))
        sam = sam.to(device="cuda")
        predictor = SamPredictor(sam)
        img = _img(600, 800, seed=5)                                        # DIOR-like, long side 800 -> 1024
        predictor.set_image(img)
        assert predictor.input_size == (768, 1024) and predictor.original_size == (600, 800)
        feats = predictor.features.clone()
        ref_in = np.array(Image.fromarray(img).resize((1024, 768), Image.BILINEAR))     # what the reference feeds its encoder
        want = sam.engine.encode(torch.from_numpy(ref_in).cuda())
        torch.cuda.synchronize()
        assert torch.equal(feats, want)
    finally:
        sys.path.remove(samrs_b200.DROPIN_PATH)
        for k in [k for k in sys.modules if k == "segment_anything" or k.startswith("segment_anything.")]:
            del sys.modules[k]

"""Rotated box -> mask prompt (`main_sam_rbox_mask_instance.py:125-141`): the device rasteriser against the driver's own
OpenCV recipe (oracle/rbox_prompt_oracle.py executes cv2 itself).  fillPoly's raster must match bit for bit - checked through
prompts of 1024 x 1024 images, where both resizes are exact (values -1000, -500, 0, 500, 1000) - and the interpolated
prompts of other image sizes to within one float32 ulp at |v| <= 1000 (6.2e-5), exact in all but isolated elements."""
import numpy as np
import pytest
import torch

from oracle import rbox_prompt_oracle as RB
from samrs_b200 import synth


def _polys(seed, n, H, W):
    rng = np.random.default_rng(seed)
    c = np.stack([rng.uniform(0.1 * W, 0.9 * W, n), rng.uniform(0.1 * H, 0.9 * H, n)], 1)
    w, h, th = rng.uniform(6, 0.5 * W, n), rng.uniform(4, 0.4 * H, n), rng.uniform(-np.pi / 2, np.pi / 2, n)
    v1 = np.stack([w / 2 * np.cos(th), w / 2 * np.sin(th)], 1)
    v2 = np.stack([-h / 2 * np.sin(th), h / 2 * np.cos(th)], 1)
    p = np.stack([c - v1 - v2, c + v1 - v2, c + v1 + v2, c - v1 + v2], 1)
    p[..., 0] = np.clip(p[..., 0], 0, W - 1)
    p[..., 1] = np.clip(p[..., 1], 0, H - 1)
    return p.astype(np.float32)


def test_oracle_recipe_shapes_and_values():
    m = RB.mask_prompts(synth.rbox_polys(3, 3), (1024, 1024))
    assert m.shape == (3, 1, 256, 256) and m.dtype == np.float32
    assert set(np.unique(m)).issubset({-1000.0, -500.0, 0.0, 500.0, 1000.0})
    m2 = RB.mask_prompts(_polys(1, 2, 600, 800), (600, 800))
    assert m2.shape == (2, 1, 256, 256) and (m2[:, :, 192:, :] == -1000.0).all()        # padded rows (600 -> 768 of 1024)


@pytest.mark.gpu
def test_device_rasteriser_matches_the_cv2_recipe():
    from samrs_b200.engine import Engine
    eng = Engine("vit_t64", "cuda:0")
    total = exact = 0
    worst = 0.0
    for seed, (H, W) in enumerate([(1024, 1024), (1024, 1024), (600, 800), (704, 1000), (1024, 768), (333, 517), (1200, 900)]):
        polys = _polys(seed, 16, H, W)
        if seed == 1:                                            # arbitrary (also self-intersecting) quadrilaterals
            rng = np.random.default_rng(99)
            polys = np.stack([rng.uniform(0, W - 1, (16, 4)), rng.uniform(0, H - 1, (16, 4))], -1).astype(np.float32)
        want = RB.mask_prompts(polys, (H, W))
        got = eng.rbox_mask_prompts(torch.from_numpy(polys).cuda(), (H, W)).cpu().numpy()
        assert got.shape == want.shape
        if (H, W) == (1024, 1024):
            assert np.array_equal(got, want), f"fillPoly raster differs for seed {seed}"
        diff = np.abs(got - want)
        worst = max(worst, float(diff.max()))
        total += want.size
        exact += int((diff == 0).sum())
        assert diff.max() <= 6.2e-5
    print(f"rbox mask prompts: {exact} of {total} elements identical to the cv2 recipe, max |diff| {worst:.2e}")
    assert exact >= 0.9999 * total
    with pytest.raises(ValueError):
        eng.rbox_mask_prompts(torch.tensor([[[5.0, 5.0], [2000.0, 5.0], [2000.0, 50.0], [5.0, 50.0]]]).cuda(), (600, 800))
    eng.close()

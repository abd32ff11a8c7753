"""Pins oracle/sam_oracle.py to the reference: the committed fixtures are outputs of the
reference's own segment_anything (oracle/make_golden.py), so agreement here means the
restatement reproduces the reference on the same seeded checkpoint and prompts."""
import os

import numpy as np
import pytest
import torch

from oracle import sam_oracle as O
from samrs_b200.config import geometry
from samrs_b200 import synth
from samrs_b200.weights import synthetic_state_dict

TINY = ["t64_box", "t80_box", "t64_pts5", "t64_point1", "t80_maskprompt", "t64_box_pts"]
_SD = {}


def _sd(variant):
    if variant not in _SD:
        _SD.clear()
        _SD[variant] = synthetic_state_dict(variant, 0)
    return _SD[variant]


def _run_case(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    variant = str(z["variant"])
    g, w = geometry(variant), _sd(variant)
    torch.set_num_threads(int(z["threads"]))
    with torch.no_grad():
        feat = O.set_image(w, g, synth.tile(int(z["tile_idx"])))
        get = lambda k: torch.from_numpy(z["prompt_" + k]) if ("prompt_" + k) in z else None
        masks, iou, low = O.predict_torch(w, g, feat, get("point_coords"), get("point_labels"),
                                          get("boxes"), get("mask_input"), bool(z["multimask"]))
    return z, feat.numpy(), masks.numpy(), iou.numpy(), low.numpy()


def _check(z, feat, masks, iou, low):
    # same ATen kernels, same thread count -> agreement to fp32 round-off
    np.testing.assert_allclose(feat[0, ::8, ::4, ::4], z["feat_sub"], rtol=0, atol=2e-5)
    assert abs(np.abs(feat.astype(np.float64)).sum() - float(z["feat_abssum"])) < 1e-4 * float(z["feat_abssum"])
    np.testing.assert_allclose(iou, z["iou"], rtol=0, atol=1e-5)
    if "low_res" in z:
        np.testing.assert_allclose(low, z["low_res"], rtol=0, atol=1e-5)
    else:
        np.testing.assert_allclose(low[:, :, ::2, ::2], z["low_sub"], rtol=0, atol=1e-5)
    pop = masks.reshape(masks.shape[0], masks.shape[1], -1).sum(-1)
    # thresholded pixels may flip where |logit| ~ 1e-7; allow a handful
    assert np.abs(pop - z["mask_popcount"]).max() <= 16


@pytest.mark.parametrize("name", TINY)
def test_oracle_matches_reference_tiny(golden_dir, name):
    _check(*_run_case(golden_dir, name))


@pytest.mark.slow
def test_oracle_matches_reference_vit_b(golden_dir):
    _check(*_run_case(golden_dir, "b_box"))


def test_painter_and_fused_epilogue_restatement(golden_dir):
    """painter_reduce == the driver's loop; upsample_threshold_paint == interpolate+threshold+painter."""
    z = np.load(os.path.join(golden_dir, "t64_box.npz"))
    low = z["low_res"]
    masks = np.unpackbits(z["mask_bits"], axis=-1).astype(bool)
    labels = z["labels"]
    assert np.array_equal(O.painter_reduce(masks, labels), z["label_map"])
    assert np.array_equal(O.upsample_threshold_paint(low[:, 0], labels), z["label_map"])
    # order matters (last box wins): reversing the visit order must change overlapping pixels
    rev = O.painter_reduce(masks[::-1], labels[::-1])
    assert (rev != z["label_map"]).any()


@pytest.mark.slow
def test_oracle_matches_reference_vit_h_benchmarked_configs(golden_dir):
    """ViT-H at the benchmarked size: the oracle reproduces the reference's embedding of tile 0 and of the 600 x 800 tile
    (resize + pad-after-normalise path) and, from them, the 32-box / cv2-mask-prompt / non-square logits and label maps."""
    zf = np.load(os.path.join(golden_dir, "h_feat.npz"))
    g, w = geometry("vit_h"), _sd("vit_h")
    torch.set_num_threads(8)
    with torch.no_grad():
        feat = O.set_image(w, g, synth.tile(0))
        np.testing.assert_allclose(feat.numpy(), zf["feat_t0"], rtol=0, atol=5e-5)
        for name in ("h_box32", "h_mask8"):
            z = np.load(os.path.join(golden_dir, name + ".npz"))
            get = lambda k: torch.from_numpy(z["prompt_" + k]) if ("prompt_" + k) in z else None
            masks, iou, low = O.predict_torch(w, g, feat, None, None, get("boxes"), get("mask_input"), False)
            np.testing.assert_allclose(low.numpy(), z["low_res"], rtol=0, atol=1e-5)
            np.testing.assert_allclose(iou.numpy(), z["iou"], rtol=0, atol=1e-5)
            lm = O.painter_reduce(masks[:, 0].numpy(), z["labels"])
            assert (lm != z["label_map"]).sum() <= 16          # |logit| ~ 1e-7 pixels
        z = np.load(os.path.join(golden_dir, "h_ns_box8.npz"))
        H, Wd = [int(v) for v in z["image_hw"]]
        img = synth.tile(int(z["tile_idx"]), 1024)[:H, :Wd].copy()
        feat = O.set_image(w, g, img)
        np.testing.assert_allclose(feat.numpy(), zf["feat_ns"], rtol=0, atol=5e-5)
        in_hw = tuple(int(v) for v in z["input_size"])
        assert O.apply_image(img).shape[:2] == in_hw
        tb = O.apply_boxes(torch.from_numpy(z["prompt_boxes"]), (H, Wd))
        masks, iou, low = O.predict_torch(w, g, feat, None, None, tb, None, False, input_size=in_hw, original_size=(H, Wd))
        np.testing.assert_allclose(low.numpy(), z["low_res"], rtol=0, atol=1e-5)
        assert masks.shape[-2:] == (H, Wd)
        assert (O.painter_reduce(masks[:, 0].numpy(), z["labels"]) != z["label_map"]).sum() <= 16


def test_rbox_mask_prompt_fixture_is_the_cv2_recipe(golden_dir):
    """The committed mask prompts of h_mask8 are what the driver's cv2 calls produce for those polygons (blended edges)."""
    from oracle import rbox_prompt_oracle as RB
    z = np.load(os.path.join(golden_dir, "h_mask8.npz"))
    m = RB.mask_prompts(z["polys"], (1024, 1024))
    assert np.array_equal(m, z["prompt_mask_input"])
    inner = np.abs(m) < 999.0
    assert inner.any() and m.min() == -1000.0 and m.max() == 1000.0

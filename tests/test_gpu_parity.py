"""End-to-end parity on a B200, through the C ABI (samrs_b200.engine -> libsamrs_b200.so).

Three anchors:
  * committed golden fixtures = outputs of the reference's own segment_anything (oracle/make_golden.py);
  * the CPU oracle (oracle/sam_oracle.py) run here on the same seeded checkpoint / tile / prompts;
  * integer work (threshold, painter label map) bit-exact given identical low-res logits.
Tolerance on mask logits: 1e-3 absolute (BASELINE.json north_star).
"""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from samrs_b200 import synth  # noqa: E402
from samrs_b200.config import geometry  # noqa: E402
from samrs_b200.weights import synthetic_state_dict  # noqa: E402

LOGIT_TOL = 1e-3
_ENG = {}


def engine_for(variant):
    from samrs_b200.engine import Engine
    if variant not in _ENG:
        for k in list(_ENG):
            _ENG.pop(k).close()
        e = Engine(variant, "cuda:0")
        e.load_state_dict(synthetic_state_dict(variant, 0))
        _ENG[variant] = e
    return _ENG[variant]


def run_engine(z):
    variant = str(z["variant"])
    eng = engine_for(variant)
    img = torch.from_numpy(synth.tile(int(z["tile_idx"]))).cuda()
    feat = eng.encode(img)
    get = lambda k: torch.from_numpy(z["prompt_" + k]).cuda() if ("prompt_" + k) in z else None
    low, iou = eng.decode(boxes=get("boxes"), point_coords=get("point_coords"), point_labels=get("point_labels"),
                          mask_input=get("mask_input"), multimask_output=bool(z["multimask"]))
    masks = eng.postprocess(low, (1024, 1024), (1024, 1024))
    torch.cuda.synchronize()
    return eng, feat.cpu().numpy(), low.cpu().numpy(), iou.cpu().numpy(), masks.cpu().numpy()


CASES = ["t64_box", "t80_box", "t64_pts5", "t64_point1", "t80_maskprompt", "t64_box_pts", "b_box", "h_box"]


@pytest.mark.parametrize("name", CASES)
def test_engine_matches_reference_golden(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    _, feat, low, iou, masks = run_engine(z)
    fsub = feat[0, ::8, ::4, ::4]
    ferr = np.abs(fsub - z["feat_sub"]).max()
    if "low_res" in z:
        lerr = np.abs(low - z["low_res"]).max()
    else:
        lerr = np.abs(low[:, :, ::2, ::2] - z["low_sub"]).max()
    ierr = np.abs(iou - z["iou"]).max()
    pop = masks.reshape(masks.shape[0], masks.shape[1], -1).sum(-1)
    flips = np.abs(pop - z["mask_popcount"])
    print(f"{name}: feat err {ferr:.3e} low-res logit err {lerr:.3e} (absmax {float(z['low_absmax']):.3f}) "
          f"iou err {ierr:.3e} popcount delta {flips.ravel().tolist()}")
    assert lerr < LOGIT_TOL
    assert ierr < LOGIT_TOL
    assert ferr < 2e-2          # features are O(1) LayerNorm outputs; fp16 operands through up to 32 blocks
    # thresholded pixels: report, and bound the damage (logits within 1e-3 of zero may flip)
    assert flips.max() <= 0.01 * 1024 * 1024


@pytest.mark.parametrize("variant,kind", [("vit_t64", "box"), ("vit_t80", "box")])
def test_engine_matches_oracle_full_tensors(variant, kind):
    """Full-tensor comparison against the CPU oracle (not subsampled), 8 boxes, tile 7."""
    from oracle import sam_oracle as O
    g, w = geometry(variant), synthetic_state_dict(variant, 0)
    eng = engine_for(variant)
    img = synth.tile(7)
    boxes = torch.from_numpy(synth.hboxes(7, 8))
    with torch.no_grad():
        f_ref = O.set_image(w, g, img)
        m_ref, i_ref, l_ref = O.predict_torch(w, g, f_ref, None, None, boxes, None, False)
    feat = eng.encode(torch.from_numpy(img).cuda())
    low, iou = eng.decode(boxes=boxes.cuda(), multimask_output=False)
    masks = eng.postprocess(low, (1024, 1024), (1024, 1024))
    torch.cuda.synchronize()
    assert (feat.cpu() - f_ref).abs().max().item() < 2e-2
    assert (low.cpu() - l_ref).abs().max().item() < LOGIT_TOL
    assert (iou.cpu() - i_ref).abs().max().item() < LOGIT_TOL
    flips = (masks.cpu() != m_ref).sum().item()
    print(f"{variant}: mask pixel flips {flips} of {m_ref.numel()}")
    assert flips < 0.005 * m_ref.numel()


def test_many_prompts_span_several_decoder_passes():
    """BASELINE config 4 density and beyond: 64 tiny boxes (exactly one decoder pass) and 70 boxes (64 + 6: the engine
    splits internally) against the oracle fed with the engine's own features; splitting must not change any prompt."""
    from oracle import sam_oracle as O
    variant = "vit_t64"
    g, w = geometry(variant), synthetic_state_dict(variant, 0)
    eng = engine_for(variant)
    img = synth.tile(9)
    feat = eng.encode(torch.from_numpy(img).cuda())
    boxes = torch.from_numpy(np.concatenate([synth.hboxes(9, 64, tiny=True), synth.hboxes(10, 6)], 0))
    low70, iou70 = eng.decode(boxes=boxes.cuda(), multimask_output=False)
    low64, iou64 = eng.decode(boxes=boxes[:64].cuda(), multimask_output=False)
    low6, iou6 = eng.decode(boxes=boxes[64:].cuda(), multimask_output=False)
    torch.cuda.synchronize()
    assert tuple(low70.shape) == (70, 1, 256, 256)
    assert torch.equal(low70[:64], low64) and torch.equal(low70[64:], low6)
    assert torch.equal(iou70[:64], iou64) and torch.equal(iou70[64:], iou6)
    with torch.no_grad():
        _, i_ref, l_ref = O.predict_torch(w, g, feat.cpu(), None, None, boxes, None, False)
    assert (low70.cpu() - l_ref).abs().max().item() < LOGIT_TOL
    assert (iou70.cpu() - i_ref).abs().max().item() < LOGIT_TOL


def test_decoder_only_from_reference_features(golden_dir):
    """Isolates the fp32 decoder: feed the ORACLE's features, compare logits (decoder runs in fp32 -> tight)."""
    from oracle import sam_oracle as O
    variant = "vit_t64"
    g, w = geometry(variant), synthetic_state_dict(variant, 0)
    eng = engine_for(variant)
    boxes = torch.from_numpy(synth.hboxes(9, 6))
    with torch.no_grad():
        f_ref = O.set_image(w, g, synth.tile(9))
        _, i_ref, l_ref = O.predict_torch(w, g, f_ref, None, None, boxes, None, True)
    eng.set_features(f_ref.cuda())
    low, iou = eng.decode(boxes=boxes.cuda(), multimask_output=True)
    torch.cuda.synchronize()
    assert (low.cpu() - l_ref).abs().max().item() < 2e-5
    assert (iou.cpu() - i_ref).abs().max().item() < 2e-5


def test_fused_epilogue_bit_exact(golden_dir):
    """Given identical low-res logits the uint8 outputs must be bit-identical to the reference's:
    masks == interpolate+threshold, label map == the driver's painter loop."""
    z = np.load(os.path.join(golden_dir, "t64_box.npz"))
    eng = engine_for("vit_t64")
    low = torch.from_numpy(z["low_res"]).cuda()
    masks = eng.postprocess(low, (1024, 1024), (1024, 1024)).cpu().numpy()
    ref_masks = np.unpackbits(z["mask_bits"], axis=-1).astype(bool)
    assert np.array_equal(masks[:, 0], ref_masks)
    canvas = torch.full((1024, 1024), 255, dtype=torch.uint8, device="cuda")
    eng.semantic_reduce(low, torch.from_numpy(z["labels"]).cuda(), canvas)
    assert np.array_equal(canvas.cpu().numpy(), z["label_map"])
    # chunked calls (20 + 12 in the driver) compose: later chunks overwrite earlier ones
    canvas2 = torch.full((1024, 1024), 255, dtype=torch.uint8, device="cuda")
    lab = torch.from_numpy(z["labels"]).cuda()
    eng.semantic_reduce(low[:3], lab[:3], canvas2)
    eng.semantic_reduce(low[3:], lab[3:], canvas2)
    assert np.array_equal(canvas2.cpu().numpy(), z["label_map"])


def test_postprocess_general_sizes_match_oracle():
    """Non-square original sizes: 256 -> 1024 -> crop -> (H,W), against torch's F.interpolate on the host."""
    from oracle import sam_oracle as O
    eng = engine_for("vit_t64")
    g = torch.Generator().manual_seed(3)
    low = torch.randn((3, 1, 256, 256), generator=g)
    for input_size, original in [((1024, 683), (1500, 1000)), ((768, 1024), (600, 800)), ((1024, 1024), (1024, 1024))]:
        ref = O.postprocess_masks(low, input_size, original)
        out = eng.postprocess(low.cuda(), input_size, original, return_logits=True).cpu()
        # the fractional source index is computed in fp32 (values up to ~1024, ulp 6e-5): the order of the
        # multiply/subtract rounding differs between ATen builds, so logits agree to ~1e-4 on unit-scale input
        assert (out - ref).abs().max().item() < 1e-4
        outm = eng.postprocess(low.cuda(), input_size, original).cpu()
        assert (outm != (ref > 0)).sum().item() <= 2e-4 * ref.numel()


def test_dropin_predictor_surface(tmp_path):
    """The drop-in `segment_anything` package: registry -> .to(cuda) -> SamPredictor -> predict_torch,
    used exactly as main_sam_hbox_semantic.py:87-89,155,174-189 does (ViT-B to keep the checkpoint small)."""
    import samrs_b200
    sys.path.insert(0, samrs_b200.DROPIN_PATH)
    try:
        for k in [k for k in sys.modules if k == "segment_anything" or k.startswith("segment_anything.")]:
            del sys.modules[k]
        from segment_anything import SamPredictor, sam_model_registry
        ckpt = tmp_path / "sam_vit_b.pth"
        torch.save(synthetic_state_dict("vit_b", 0), ckpt)
        sam = sam_model_registry["vit_b"](checkpoint=str(ckpt))
        sam = sam.to(device="cuda")
        predictor = SamPredictor(sam)
        with pytest.raises(RuntimeError):
            predictor.predict_torch(None, None, boxes=torch.zeros(1, 4, device="cuda"))
        img = synth.tile(0)
        predictor.set_image(img)
        boxes = torch.from_numpy(synth.hboxes(0, 2)).cuda()
        tb = predictor.transform.apply_boxes_torch(boxes, img.shape[:2])
        masks, iou, low = predictor.predict_torch(point_coords=None, point_labels=None, boxes=tb, mask_input=None,
                                                  multimask_output=False)
        assert masks.dtype == torch.bool and tuple(masks.shape) == (2, 1, 1024, 1024) and masks.is_cuda
        assert tuple(iou.shape) == (2, 1) and tuple(low.shape) == (2, 1, 256, 256)
        z = np.load(os.path.join(os.path.dirname(__file__), "golden", "b_box.npz"))
        assert np.abs(low.cpu().numpy()[:, :, ::2, ::2] - z["low_sub"]).max() < LOGIT_TOL
        sam_masks = masks.squeeze(1).cpu().numpy()
        assert sam_masks.shape == (2, 1024, 1024)
        # instance branch (main_sam_hbox_semantic.py:200-204): fused on-device RLE == encoding the driver's host masks
        from oracle import rle_oracle
        from samrs_b200 import rle as host_rle
        from samrs_b200.stream import instance_tile
        recs = instance_tile(predictor, sam.engine, img, boxes, [3, 5], categories=[str(i) for i in range(18)], chunk=20)
        assert len(recs) == 2 and recs[1]["label"] == 5 and recs[1]["category"] == "5"
        for j, r in enumerate(recs):
            assert r["size"] == int(sam_masks[j].sum())
            assert host_rle.coco_string_decode(r["mask"]["counts"]) == rle_oracle.mask_to_rle(sam_masks[j])
    finally:
        sys.path.remove(samrs_b200.DROPIN_PATH)
        for k in [k for k in sys.modules if k == "segment_anything" or k.startswith("segment_anything.")]:
            del sys.modules[k]

"""Generate tests/golden/*.npz by running the REFERENCE's own segment_anything.

Run in the build container only (it needs /root/reference, which does not
exist on the GPU box):

    python oracle/make_golden.py            # all cases
    python oracle/make_golden.py t64_box    # one case

For every case the unmodified reference package is imported from
`/root/reference/Generate Dataset`, built through its own `_build_sam`
(`segment_anything/build_sam.py:55-107`), loaded (strict) with the seeded
synthetic checkpoint of `samrs_b200.weights`, and driven through
`SamPredictor.set_image` / `predict_torch` exactly as
`main_sam_hbox_semantic.py:155,174-181` does.  Outputs are stored subsampled
where full tensors would bloat the repo; checksums cover the full tensors.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference/Generate Dataset"

from samrs_b200 import synth  # noqa: E402
from samrs_b200.config import geometry  # noqa: E402
from samrs_b200.weights import synthetic_state_dict  # noqa: E402

# name -> (variant, tile idx, prompt kind, n prompts, multimask, store full low-res?)
CASES = {
    "t64_box": ("vit_t64", 0, "box", 4, False, True),
    "t80_box": ("vit_t80", 1, "box", 3, False, False),
    "t64_pts5": ("vit_t64", 2, "pts5", 3, False, False),
    "t64_point1": ("vit_t64", 3, "point1", 3, True, False),
    "t80_maskprompt": ("vit_t80", 4, "mask", 2, False, False),
    "t64_box_pts": ("vit_t64", 5, "box+pts5", 2, True, False),
    "b_box": ("vit_b", 0, "box", 2, False, False),
    "h_box": ("vit_h", 0, "box", 2, False, False),
}


def prompts_for(kind: str, idx: int, n: int):
    """-> dict of numpy prompt arrays in the 1024 input frame (tiles are 1024^2, so
    apply_boxes_torch is the identity scale)."""
    out = {}
    if "box" in kind:
        out["boxes"] = synth.hboxes(idx, n)
    if "pts5" in kind:
        out["point_coords"] = synth.rboxes_5pt(idx, n)
        out["point_labels"] = np.ones((n, 5), dtype=np.int32)
    if kind == "point1":
        b = synth.hboxes(idx, n)
        out["point_coords"] = ((b[:, :2] + b[:, 2:]) / 2)[:, None, :].astype(np.float32)
        out["point_labels"] = np.ones((n, 1), dtype=np.int32)
    if kind == "mask":
        out["mask_input"] = synth.mask_prompts(idx, n)
    return out


def run_reference(variant: str, sd, img: np.ndarray, pr: dict, multimask: bool):
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from segment_anything import SamPredictor  # the reference's package
    from segment_anything.build_sam import _build_sam

    g = geometry(variant)
    sam = _build_sam(g.embed_dim, g.depth, g.num_heads, list(g.global_attn_indexes))
    sam.load_state_dict(sd, strict=True)
    pred = SamPredictor(sam)
    pred.set_image(img)
    t = {k: torch.from_numpy(v) for k, v in pr.items()}
    boxes = t.get("boxes")
    if boxes is not None:
        boxes = pred.transform.apply_boxes_torch(boxes, img.shape[:2])
    masks, iou, low = pred.predict_torch(
        point_coords=t.get("point_coords"), point_labels=t.get("point_labels"),
        boxes=boxes, mask_input=t.get("mask_input"), multimask_output=multimask)
    return pred.features, masks, iou, low


def main() -> None:
    torch.set_num_threads(8)
    torch.manual_seed(0)
    outdir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(outdir, exist_ok=True)
    want = sys.argv[1:] or list(CASES)
    cache = {}
    for name in want:
        variant, idx, kind, n, multimask, full = CASES[name]
        if variant not in cache:
            cache = {variant: synthetic_state_dict(variant, seed=0)}
        sd = cache[variant]
        img = synth.tile(idx)
        pr = prompts_for(kind, idx, n)
        lab = synth.labels(idx, n)
        with torch.no_grad():
            feat, masks, iou, low = run_reference(variant, sd, img, pr, multimask)
        feat, masks, iou, low = feat.numpy(), masks.numpy(), iou.numpy(), low.numpy()
        rec = {
            "variant": variant, "tile_idx": idx, "kind": kind, "multimask": multimask,
            "labels": lab,
            "feat_sub": feat[0, ::8, ::4, ::4].copy(),
            "feat_sum": np.float64(feat.astype(np.float64).sum()),
            "feat_abssum": np.float64(np.abs(feat.astype(np.float64)).sum()),
            "iou": iou,
            "low_sum": low.astype(np.float64).sum(axis=(2, 3)),
            "low_abssum": np.abs(low.astype(np.float64)).sum(axis=(2, 3)),
            "low_absmax": np.float64(np.abs(low).max()),
            "mask_popcount": masks.reshape(masks.shape[0], masks.shape[1], -1).sum(-1).astype(np.int64),
            "torch_version": torch.__version__, "threads": torch.get_num_threads(),
        }
        rec.update({"prompt_" + k: v for k, v in pr.items()})
        if full:
            rec["low_res"] = low
            # the driver's painter reduce (main_sam_hbox_semantic.py:162,195-199), restated inline
            seg = np.full(img.shape[:2], 255, dtype=np.uint8)
            for j in range(masks.shape[0]):
                r, c = np.nonzero(masks[j, 0])
                seg[r, c] = lab[j]
            rec["label_map"] = seg
            rec["mask_bits"] = np.packbits(masks[:, 0], axis=-1)
        else:
            rec["low_sub"] = low[:, :, ::2, ::2].copy()
        path = os.path.join(outdir, name + ".npz")
        np.savez_compressed(path, **rec)
        print(f"{name}: feat|sum|={rec['feat_abssum']:.3f} low absmax={rec['low_absmax']:.4f} "
              f"popcount={rec['mask_popcount'].ravel().tolist()} -> {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()

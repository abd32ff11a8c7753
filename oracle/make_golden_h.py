"""Generate the ViT-H fixtures of the BENCHMARKED configurations by running the REFERENCE's own segment_anything.

Build container only (needs /root/reference):   python oracle/make_golden_h.py [case ...]

One reference `Sam` (ViT-H, seeded synthetic checkpoint) is built once; `set_image` runs once per tile and every
prompt case decodes against that embedding exactly as the drivers do (`main_sam_hbox_semantic.py:155-189`,
`main_sam_rbox_mask_instance.py:125-164`).  Stored per case (tests/golden/<case>.npz):

  * prompts, labels, the FULL fp32 low-res logits `low_res` (B,1,256,256) and IoU predictions;
  * `label_map`: the driver's painter reduce of the reference's bool masks (uint8, 255 = untouched);
  * `mask_popcount`, `mask_crc`: per-mask true-pixel counts and CRC32 of the packed full-resolution bits, so a test can
    check that `oracle.postprocess_masks(low_res) > 0` re-creates the reference's masks before counting flips;
  * h_box32 also holds the driver's 20 + 12 chunking (`low_res_chunked` differences, label map of the chunked run).
`h_feat.npz` holds the reference image embedding of tile 0 (fp32, full) and of the 600x800 tile.
"""
from __future__ import annotations

import os
import sys
import time
import zlib

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference/Generate Dataset"

from oracle import rbox_prompt_oracle as RB  # noqa: E402
from samrs_b200 import synth  # noqa: E402
from samrs_b200.config import geometry  # noqa: E402
from samrs_b200.weights import scale_logits_, synthetic_state_dict  # noqa: E402

VARIANT = "vit_h"
LOGIT_SCALE = 32.0
NS_TILE, NS_HW = 21, (600, 800)          # the non-square case: tile seed and (H, W)


def crc_masks(masks: np.ndarray) -> np.ndarray:
    return np.array([zlib.crc32(np.packbits(m).tobytes()) for m in masks], dtype=np.int64)


def painter(masks: np.ndarray, labels) -> np.ndarray:
    """main_sam_hbox_semantic.py:162,195-199 restated inline (last box wins)."""
    seg = np.full(masks.shape[-2:], 255, dtype=np.uint8)
    for j in range(masks.shape[0]):
        r, c = np.nonzero(masks[j])
        seg[r, c] = labels[j]
    return seg


def predict(pred, img_hw, pr, chunk=None):
    t = {k: torch.from_numpy(v) for k, v in pr.items()}
    n = next(iter(t.values())).shape[0]
    outs = []
    step = chunk or n
    for s in range(0, n, step):
        boxes = t.get("boxes")
        if boxes is not None:
            boxes = pred.transform.apply_boxes_torch(boxes[s:s + step], img_hw)
        pc, pl, mi = t.get("point_coords"), t.get("point_labels"), t.get("mask_input")
        if pc is not None:
            pc = pred.transform.apply_coords_torch(pc[s:s + step], img_hw)
        outs.append(pred.predict_torch(point_coords=pc, point_labels=None if pl is None else pl[s:s + step], boxes=boxes,
                                       mask_input=None if mi is None else mi[s:s + step], multimask_output=False))
    return [torch.cat([o[i] for o in outs], 0).numpy() for i in range(3)]


def record(name, pr, labels, masks, iou, low, extra=None):
    rec = {"variant": VARIANT, "labels": labels, "low_res": low, "iou": iou, "low_absmax": np.float64(np.abs(low).max()),
           "label_map": painter(masks[:, 0], labels), "mask_popcount": masks[:, 0].reshape(masks.shape[0], -1).sum(-1).astype(np.int64),
           "mask_crc": crc_masks(masks[:, 0]), "torch_version": torch.__version__, "threads": torch.get_num_threads()}
    rec.update({"prompt_" + k: v for k, v in pr.items()})
    rec.update(extra or {})
    path = os.path.join(ROOT, "tests", "golden", name + ".npz")
    np.savez_compressed(path, **rec)
    print(f"{name}: B={low.shape[0]} low absmax={rec['low_absmax']:.4f} labelled px={(rec['label_map'] != 255).sum()} "
          f"-> {os.path.getsize(path) / 2**20:.1f} MiB", flush=True)


def main() -> None:
    torch.set_num_threads(8)
    want = set(sys.argv[1:])
    on = lambda c: not want or c in want
    sys.path.insert(0, REF)
    from segment_anything import SamPredictor          # the reference's package
    from segment_anything.build_sam import _build_sam
    g = geometry(VARIANT)
    sd = synthetic_state_dict(VARIANT, 0)
    sam = _build_sam(g.embed_dim, g.depth, g.num_heads, list(g.global_attn_indexes))
    sam.load_state_dict(sd, strict=True)
    pred = SamPredictor(sam)
    feats = {}
    with torch.no_grad():
        # ---------------------------------------------------------------- tile 0, 1024 x 1024
        img = synth.tile(0)
        t0 = time.time()
        pred.set_image(img)
        print(f"reference set_image (ViT-H, 8 threads): {time.time() - t0:.1f} s", flush=True)
        feats["feat_t0"] = pred.features.numpy().copy()
        hw = img.shape[:2]
        lab32, lab64 = synth.labels(0, 32), synth.labels(0, 64)
        if on("h_box32"):
            pr = {"boxes": synth.hboxes(0, 32)}
            masks, iou, low = predict(pred, hw, pr)
            mc, ic, lc = predict(pred, hw, pr, chunk=20)          # the driver's 20 + 12
            extra = {"chunk": 20, "low_chunked_maxdiff": np.float64(np.abs(lc - low).max()),
                     "label_map_chunked": painter(mc[:, 0], lab32), "mask_crc_chunked": crc_masks(mc[:, 0]),
                     "mask_popcount_chunked": mc[:, 0].reshape(32, -1).sum(-1).astype(np.int64),
                     "low_res_chunked_delta": (lc - low).astype(np.float32)}
            record("h_box32", pr, lab32, masks, iou, low, extra)
        if on("h_pts5_32"):
            pr = {"point_coords": synth.rboxes_5pt(0, 32), "point_labels": np.ones((32, 5), dtype=np.int32)}
            record("h_pts5_32", pr, lab32, *predict(pred, hw, pr))
        if on("h_mask8"):
            polys = synth.rbox_polys(0, 8)
            pr = {"mask_input": RB.mask_prompts(polys, hw)}
            record("h_mask8", pr, lab32[:8], *predict(pred, hw, pr), extra={"polys": polys})
        if on("h_tiny64"):
            pr = {"boxes": synth.hboxes(0, 64, tiny=True)}
            record("h_tiny64", pr, lab64, *predict(pred, hw, pr))
        if on("h_box32_s32"):
            # SAM-like logit magnitude: same embedding, hyper-network output layers x 32
            ssd = {k: v.clone() for k, v in sam.state_dict().items()}
            scale_logits_(ssd, LOGIT_SCALE)
            sam.load_state_dict(ssd, strict=True)
            pr = {"boxes": synth.hboxes(0, 32)}
            record("h_box32_s32", pr, lab32, *predict(pred, hw, pr), extra={"logit_scale": np.float64(LOGIT_SCALE)})
            sam.load_state_dict(sd, strict=True)
        # ---------------------------------------------------------------- 600 x 800 tile: resize to 768 x 1024, pad, crop back
        if on("h_ns_box8"):
            img = synth.tile(NS_TILE, 1024)[: NS_HW[0], : NS_HW[1]].copy()
            pred.set_image(img)
            feats["feat_ns"] = pred.features.numpy().copy()
            b = synth.hboxes(NS_TILE, 8, size=600)
            pr = {"boxes": b}
            masks, iou, low = predict(pred, img.shape[:2], pr)
            record("h_ns_box8", pr, synth.labels(NS_TILE, 8), masks, iou, low,
                   extra={"tile_idx": NS_TILE, "image_hw": np.array(NS_HW), "input_size": np.array(pred.input_size)})
    if not want or "h_feat" in want or len(feats) == 2:
        path = os.path.join(ROOT, "tests", "golden", "h_feat.npz")
        np.savez_compressed(path, **feats)
        print(f"h_feat: {list(feats)} -> {os.path.getsize(path) / 2**20:.1f} MiB")


if __name__ == "__main__":
    main()

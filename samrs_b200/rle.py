"""Host side of the instance payload: turn the device encoder's (counts, offsets, area) into the reference's records.

The device part is `Engine.rle_encode` (`samrs_rle_encode`, csrc/rle.cuh).  Here:

* `to_rle_dicts`   -> `[{"size": [h, w], "counts": [...]}, ...]`, the uncompressed form `mask_to_rle_pytorch` returns
                      (`segment_anything/utils/amg.py:107-135`), which `rle_to_mask` (`amg.py:138-149`) inverts;
* `rle_to_mask`    -> the inverse, for round trips;
* `coco_string` / `coco_string_decode` -> the compressed ASCII form stored in the drivers' `ins/*.pkl`
                      (`main_sam_hbox_semantic.py:200-201`).  That string is produced by pycocotools, which is neither
                      vendored nor installed: this is a restatement of its published `rleToString` /
                      `rleFrString` scheme and its **parity is unpinned** - tests only check the round trip;
* `instance_records` -> the per-instance dicts `{mask, bbox, category, label, size}` the drivers append
                      (`main_sam_hbox_semantic.py:202-204`).
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Sequence

import numpy as np
import torch


def to_rle_dicts(counts: torch.Tensor, offsets: torch.Tensor, h: int, w: int) -> List[Dict[str, Any]]:
    off = offsets.detach().cpu().numpy().astype(np.int64)
    if off[-1] > counts.numel():
        raise ValueError(f"rle: {int(off[-1])} runs do not fit the capacity of {counts.numel()}; encode again with a larger capacity")
    cnt = counts[: int(off[-1])].detach().cpu().numpy().astype(np.int64)
    return [{"size": [int(h), int(w)], "counts": cnt[off[b]:off[b + 1]].tolist()} for b in range(len(off) - 1)]


def rle_to_mask(rle: Dict[str, Any]) -> np.ndarray:
    h, w = rle["size"]
    counts = np.asarray(rle["counts"], dtype=np.int64)
    if counts.sum() != h * w:
        raise ValueError("rle: run lengths do not sum to h*w")
    vals = (np.arange(len(counts)) & 1).astype(bool)
    return np.repeat(vals, counts).reshape(w, h).T


def mask_to_counts(mask: np.ndarray) -> List[int]:
    """Uncompressed COCO counts of one (h, w) mask that is already on the host (column-major runs, zeros first):
    what `maskUtils.encode(np.asfortranarray(mask))` run-length encodes (main_sam_hbox_semantic.py:200).  Used by the
    harness' `pycocotools.mask` stand-in for masks a driver has copied back itself; the device path is `Engine.rle_encode`."""
    m = np.asarray(mask)
    if m.ndim != 2:
        raise ValueError("mask_to_counts expects one (h, w) mask")
    flat = (m != 0).T.reshape(-1)
    if flat.size == 0:
        return []
    edges = np.flatnonzero(flat[1:] != flat[:-1]) + 1
    idx = np.concatenate(([0], edges, [flat.size]))
    runs = np.diff(idx).tolist()
    return ([0] + runs) if flat[0] else runs


def coco_string(counts: Sequence[int]) -> str:
    """Compressed run string: each run (from the third on, its difference to the run two before) in 5-bit groups,
    little-endian, bit 5 = continuation, offset by 48 into printable ASCII.  Parity with pycocotools unpinned."""
    out = []
    for i, c in enumerate(counts):
        x = int(c)
        if i > 2:
            x -= int(counts[i - 2])
        more = True
        while more:
            ch = x & 0x1F
            x >>= 5                       # arithmetic shift: negative differences sign-extend
            more = (x != -1) if (ch & 0x10) else (x != 0)
            if more:
                ch |= 0x20
            out.append(chr(ch + 48))
    return "".join(out)


def coco_string_decode(s: str) -> List[int]:
    counts: List[int] = []
    p = 0
    while p < len(s):
        x, k, more = 0, 0, True
        while more:
            ch = ord(s[p]) - 48
            x |= (ch & 0x1F) << (5 * k)
            more = bool(ch & 0x20)
            p += 1
            k += 1
            if not more and (ch & 0x10):
                x |= -1 << (5 * k)
        if len(counts) > 2:
            x += counts[-2]
        counts.append(x)
    return counts


def instance_records(counts: torch.Tensor, offsets: torch.Tensor, area: torch.Tensor, h: int, w: int, boxes: np.ndarray,
                     labels: Sequence[int], categories: Optional[Sequence[str]] = None, compressed: bool = True,
                     rboxes: Optional[np.ndarray] = None, strings: Optional[Sequence[str]] = None) -> List[Dict[str, Any]]:
    """One dict per mask in the drivers' schema; `mask` is a COCO RLE dict (compressed string unless compressed=False).

    Without `rboxes`: `{mask, bbox, category, label, size}` (main_sam_hbox_semantic.py:204).  With `rboxes` (B,4,2), the
    rotated-box drivers' variant `{mask, rbox, rhbox, category, label, size}` (main_sam_rhbox_semantic.py:203-209), where
    `boxes` are the polygons' enclosing horizontal boxes the masks were prompted with (`rhbox`, :120-130).
    `strings`: the compressed strings when they were already produced on the device (`Engine.rle_strings`); `counts` may
    then be None."""
    sizes = area.detach().cpu().numpy()
    if strings is not None:
        rles = [{"size": [int(h), int(w)], "counts": None} for _ in strings]
    else:
        rles = to_rle_dicts(counts, offsets, h, w)
    recs = []
    for j, r in enumerate(rles):
        if strings is not None:
            m = {"size": r["size"], "counts": strings[j]}
        else:
            m = {"size": r["size"], "counts": coco_string(r["counts"]) if compressed else r["counts"]}
        rec = {"mask": m}
        if rboxes is None:
            rec["bbox"] = np.asarray(boxes[j])
        else:
            rec["rbox"], rec["rhbox"] = np.asarray(rboxes[j]), np.asarray(boxes[j])
        rec.update({"category": categories[int(labels[j])] if categories is not None else None, "label": int(labels[j]),
                    "size": int(sizes[j])})
        recs.append(rec)
    return recs


def strings_from_device(chars: torch.Tensor, char_offsets: torch.Tensor) -> List[str]:
    """Per-mask compressed strings from `Engine.rle_strings` output (host copies of the used part only)."""
    off = char_offsets.detach().cpu().numpy().astype(np.int64)
    if off[-1] > chars.numel():
        raise ValueError(f"rle: {int(off[-1])} characters do not fit the capacity of {chars.numel()}")
    buf = chars[: int(off[-1])].detach().cpu().numpy().tobytes()
    return [buf[off[b]:off[b + 1]].decode("ascii") for b in range(len(off) - 1)]


def enclosing_hboxes(rboxes: np.ndarray) -> np.ndarray:
    """(B,4,2) polygons -> (B,4) xyxy boxes, the `gt_rhboxes` of main_sam_rhbox_semantic.py:120-130 (dtype follows the input)."""
    r = np.asarray(rboxes)
    return np.stack([r[:, :, 0].min(1), r[:, :, 1].min(1), r[:, :, 0].max(1), r[:, :, 1].max(1)], axis=1)

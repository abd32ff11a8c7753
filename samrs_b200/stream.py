"""Multi-GPU tile stream: one process per GPU, tiles sharded `files[rank::world]`, weights broadcast once.

The reference's generation drivers are single-process (`for file in files`, `Generate Dataset/
main_sam_hbox_semantic.py:110`); tiles are independent, so the stream shards with no data-path collective
(SURVEY.md 8e).  The only collective is one broadcast of the packed checkpoint at start-up so that a single
rank has to read (or, here, synthesise) it.  Everything in this module is backend-agnostic: NCCL on GPUs,
gloo in the CPU tests.
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Sequence, Tuple

import numpy as np
import torch

from .config import SamGeometry
from .weights import state_dict_spec


def shard_indices(n_items: int, rank: int, world: int) -> List[int]:
    """Round-robin shard of a tile list: rank r processes items r, r+world, ...  (== files[rank::world])."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    return list(range(rank, n_items, world))


def packed_numel(g: SamGeometry) -> int:
    return sum(int(np.prod(shape)) for _, shape, _ in state_dict_spec(g))


def pack_state_dict(g: SamGeometry, sd: Dict[str, torch.Tensor], out: torch.Tensor) -> torch.Tensor:
    """Flatten a reference-layout state dict into one fp32 blob in `state_dict_spec` order."""
    off = 0
    for key, shape, _ in state_dict_spec(g):
        n = int(np.prod(shape))
        out[off:off + n].copy_(sd[key].reshape(-1).to(torch.float32))
        off += n
    return out


def unpack_state_dict(g: SamGeometry, flat: torch.Tensor) -> Dict[str, torch.Tensor]:
    """Views into the blob, keyed and shaped like `Sam.state_dict()`."""
    out, off = {}, 0
    for key, shape, _ in state_dict_spec(g):
        n = int(np.prod(shape))
        out[key] = flat[off:off + n].view(*shape)
        off += n
    return out


def broadcast_state_dict(g: SamGeometry, sd_on_src, device, src: int = 0, group=None) -> Dict[str, torch.Tensor]:
    """Collective: rank `src` supplies `sd_on_src` (a state dict, or a callable returning one so that only that
    rank pays for loading it); every rank returns the same state dict as views of one broadcast blob."""
    import torch.distributed as dist
    flat = torch.empty(packed_numel(g), dtype=torch.float32, device=device)
    if dist.get_rank(group) == src:
        sd = sd_on_src() if callable(sd_on_src) else sd_on_src
        pack_state_dict(g, sd, flat)
    dist.broadcast(flat, src=src, group=group)
    return unpack_state_dict(g, flat)


def max_over_ranks(value: float, device, group=None) -> float:
    """Timing rule for every multi-GPU number: the job takes as long as its slowest rank."""
    import torch.distributed as dist
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def semantic_tile(predictor, engine, image: np.ndarray, boxes: torch.Tensor, labels: torch.Tensor, canvas: torch.Tensor,
                  chunk: int = 20) -> torch.Tensor:
    """One iteration of the driver's per-image loop (`main_sam_hbox_semantic.py:155-199`) with the painter reduce
    fused on the device: set_image, predict_torch in chunks of `chunk` boxes, label map into `canvas`."""
    predictor.set_image(image)
    canvas.fill_(255)
    for s in range(0, boxes.shape[0], chunk):
        tb = predictor.transform.apply_boxes_torch(boxes[s:s + chunk], image.shape[:2])
        _, _, low = predictor.predict_torch(None, None, boxes=tb, mask_input=None, multimask_output=False)
        engine.semantic_reduce(low, labels[s:s + chunk], canvas)
    return canvas


def instance_tile(predictor, engine, image: np.ndarray, boxes: torch.Tensor, labels, categories=None, chunk: int = 20):
    """The instance branch of the same loop (`main_sam_hbox_semantic.py:183-204`): per box a COCO-RLE mask, its area and the
    annotation fields.  The masks never leave the device: every chunk's low-res logits go through `Engine.rle_encode`
    (fused upsample + threshold + run-length encoding), and only the runs (tens of KiB per tile) are copied back."""
    from . import rle as host_rle
    if image.shape[0] != 1024 or image.shape[1] != 1024:
        raise ValueError("instance_tile: the fused low-res path needs a 1024x1024 tile")
    predictor.set_image(image)
    records = []
    for s in range(0, boxes.shape[0], chunk):
        tb = predictor.transform.apply_boxes_torch(boxes[s:s + chunk], image.shape[:2])
        _, _, low = predictor.predict_torch(None, None, boxes=tb, mask_input=None, multimask_output=False)
        n = low.shape[0]
        cap = n * 16384
        while True:
            counts, offsets, area = engine.rle_encode(low_res=low, capacity=cap)
            total = int(offsets[-1])                     # synchronises: the host needs the payload anyway
            if total <= cap:
                break
            cap = total
        records += host_rle.instance_records(counts, offsets, area, 1024, 1024, boxes[s:s + chunk].detach().cpu().numpy(),
                                             [int(v) for v in labels[s:s + chunk]], categories)
    return records

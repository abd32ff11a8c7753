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


def _decode_chunks(predictor, engine, boxes_dev: torch.Tensor, original_hw, chunk: int):
    """predict_torch's prompt path without the full-resolution bool masks: transformed boxes -> low-res logits per chunk."""
    for s in range(0, boxes_dev.shape[0], chunk):
        tb = predictor.transform.apply_boxes_torch(boxes_dev[s:s + chunk], original_hw)
        low, _ = engine.decode(boxes=tb, multimask_output=False)
        yield s, low


def tile_outputs(predictor, engine, image: np.ndarray, boxes: torch.Tensor, labels: torch.Tensor, canvas: torch.Tensor,
                 chunk: int = 20, instance: bool = True, capacity_per_mask: int = 1 << 17):
    """Both products of one iteration of the driver's per-image loop (`main_sam_hbox_semantic.py:155-204`) from ONE
    encode: the label map (painter reduce) in `canvas` and, if `instance`, the run-length payload of every mask.

    Everything is enqueued on the current stream; nothing is copied to the host.  Returns `(canvas, payload)` with
    `payload = [(counts, offsets, area, n_masks, capacity, source, chars, char_offsets), ...]` per chunk (CUDA tensors;
    `source` = what was encoded, kept so that `fetch_runs` can encode again should the runs not fit; `chars` = the
    compressed COCO strings, also produced on the device).  1024 x 1024 tiles use the
    fused kernels (upsample + threshold + paint / + RLE straight from the 256 x 256 logits); other sizes go through the
    general postprocess and paint / encode the bool masks."""
    predictor.set_image(image)
    H, W = int(image.shape[0]), int(image.shape[1])
    fused = (H, W) == (1024, 1024)
    canvas.fill_(255)
    payload = []
    bdev = boxes.to(predictor.device, non_blocking=True)
    ldev = labels.to(device=predictor.device, dtype=torch.int32, non_blocking=True)
    for s, low in _decode_chunks(predictor, engine, bdev, (H, W), chunk):
        n = low.shape[0]
        if fused:
            engine.semantic_reduce(low, ldev[s:s + n], canvas)
            if instance:
                cap = n * capacity_per_mask
                payload.append(_encode(engine, n, cap, "low_res", low))
        else:
            masks = engine.postprocess(low, predictor.input_size, (H, W))
            engine.paint_masks(masks, ldev[s:s + n], canvas)
            if instance:
                cap = n * capacity_per_mask
                payload.append(_encode(engine, n, cap, "masks", masks))
    return canvas, payload


def semantic_tile(predictor, engine, image: np.ndarray, boxes: torch.Tensor, labels: torch.Tensor, canvas: torch.Tensor,
                  chunk: int = 20) -> torch.Tensor:
    """The semantic branch alone: set_image, chunks of `chunk` boxes, label map into `canvas`."""
    return tile_outputs(predictor, engine, image, boxes, labels, canvas, chunk, instance=False)[0]


def instance_tile(predictor, engine, image: np.ndarray, boxes: torch.Tensor, labels, categories=None, chunk: int = 20):
    """The instance branch alone (`main_sam_hbox_semantic.py:183-204`): per box a COCO-RLE mask, its area and the
    annotation fields.  The masks never leave the device; only the runs (tens of KiB per tile) are copied back."""
    from . import rle as host_rle
    H, W = int(image.shape[0]), int(image.shape[1])
    canvas = torch.empty((H, W), dtype=torch.uint8, device=predictor.device)
    lab_t = torch.as_tensor([int(v) for v in labels], dtype=torch.int32)
    _, payload = tile_outputs(predictor, engine, image, boxes, lab_t, canvas, chunk, instance=True)
    return _records_from_payload(engine, payload, H, W, boxes.detach().cpu().numpy(), [int(v) for v in labels], categories, None)


def _encode(engine, n, cap, kind, src):
    counts, offsets, area = engine.rle_encode(capacity=cap, **{kind: src})
    chars, coff = engine.rle_strings(counts, offsets, char_capacity=2 * cap + 64 * n)
    return (counts, offsets, area, n, cap, (kind, src), chars, coff)


def fetch_runs(engine, entry, reencode: bool = True):
    """`(area on the host, chars, char_offsets)` of one payload entry, after checking that runs and characters fit their
    buffers; if not, the chunk is encoded again with the exact size (noise-like masks - e.g. from the synthetic checkpoint -
    can have 100k+ runs each).  The run lengths themselves never travel: only the compressed strings are copied."""
    counts, offsets, area, n, cap, (kind, src), chars, coff = entry
    off = offsets.cpu()
    total = int(off[-1])
    if total > cap or int(coff[-1]) > chars.numel():
        if not reencode:                                  # a second thread must not drive the engine: the caller sizes the buffers
            raise RuntimeError(f"rle: {total} runs / {int(coff[-1])} characters exceed the capacity ({cap} runs) for {n} masks; "
                               "raise capacity_per_mask")
        counts, offsets, area = engine.rle_encode(capacity=total, **{kind: src})
        chars, coff = engine.rle_strings(counts, offsets, char_capacity=7 * total + 64)
    return area.cpu(), chars, coff


def _records_from_payload(engine, payload, H, W, boxes_np, labels, categories, rboxes_np, reencode: bool = True):
    from . import rle as host_rle
    records, s = [], 0
    for entry in payload:
        n = entry[3]
        area, chars, coff = fetch_runs(engine, entry, reencode)
        records += host_rle.instance_records(None, None, area, H, W, boxes_np[s:s + n], labels[s:s + n], categories,
                                             rboxes=None if rboxes_np is None else rboxes_np[s:s + n],
                                             strings=host_rle.strings_from_device(chars, coff))
        s += n
    return records


# ------------------------------------------------------------------------------------------------ the tile stream
class TileJob:
    """One image of a driver's `for file in files` loop: name, pixels (HWC uint8 array, or a callable returning one so the
    decode happens on a loader thread), horizontal boxes in original-image pixels, class ids, optional rotated boxes."""

    __slots__ = ("name", "image", "boxes", "labels", "rboxes")

    def __init__(self, name, image, boxes, labels, rboxes=None):
        self.name, self.image, self.boxes, self.labels, self.rboxes = name, image, boxes, labels, rboxes


def run(predictor, jobs: Iterable, save_dir: str = None, mapping=None, categories: Sequence[str] = None, chunk: int = 20,
        instance: bool = True, writer_threads: int = 4, loader_threads: int = 2, depth: int = 4, on_tile=None,
        capacity_per_mask: int = 1 << 19) -> Dict[str, float]:
    """The per-rank loop that replaces `main_sam_hbox_semantic.py:110-216` (and the rhbox variant when jobs carry rboxes).

      loader threads  materialise each job's image into a pinned ring slot (JPEG / PNG decode off the GPU thread)
      this thread     H2D + encode ONCE + decode in the driver's chunks + painter + RLE, all on the current stream, then
                      asynchronous D2H of the label map and the runs into the slot's pinned buffers and an event
      finisher thread waits for the slot's event, builds the instance records and hands the tile to
      writer threads  `writers.save_tile` (gray / color PNG + pickle), the reference's on-disk contract
    Slots are recycled in order, so at most `depth` tiles are in flight and the GPU thread never waits for a writer
    unless the writers fall `depth` tiles behind.  Returns counters: `tiles`, `masks`, `seconds` (wall clock until the last
    file is written), `seconds_before_writers_drain` (until every tile's outputs are on the host) and `writer_cpu_seconds`
    (time spent inside PNG / pickle encoding, summed over the writer threads)."""
    import queue
    import threading
    import time
    from concurrent.futures import ThreadPoolExecutor

    from . import writers
    engine = predictor.model._require_engine()
    device = predictor.device
    jobs = iter(jobs)
    ready: "queue.Queue" = queue.Queue(maxsize=depth)
    free_slots: "queue.Queue" = queue.Queue()
    done_q: "queue.Queue" = queue.Queue()
    errors: List[BaseException] = []
    lock = threading.Lock()
    side = torch.cuda.Stream(device=device)                    # the finisher's copies

    class Slot:
        def __init__(self):
            self.img = None                                   # pinned uint8 HWC, grown on demand
            self.canvas_h = None
            self.canvas_d = None
            self.event = torch.cuda.Event()
            self.host_payload = []

    for _ in range(depth):
        free_slots.put(Slot())

    def loader():
        try:
            while True:
                with lock:
                    job = next(jobs, None)
                if job is None:
                    break
                img = job.image() if callable(job.image) else job.image
                img = np.ascontiguousarray(img)
                slot = free_slots.get()
                if slot.img is None or slot.img.numel() < img.size:
                    slot.img = torch.empty(img.size, dtype=torch.uint8).pin_memory()
                view = slot.img[: img.size].view(*img.shape)
                view.numpy()[...] = img
                ready.put((job, slot, view))
        except BaseException as ex:                            # surfaced by the GPU thread
            errors.append(ex)
        finally:
            ready.put(None)

    def finisher(pool):
        while True:
            item = done_q.get()
            if item is None:
                break
            job, slot, H, W, payload_h = item
            try:
                slot.event.synchronize()
                label_map = slot.canvas_h[: H * W].view(H, W).numpy().copy()
                records = []
                if instance:
                    rb = None if job.rboxes is None else np.asarray(job.rboxes)
                    with torch.cuda.stream(side):              # D2H of the runs that exist, off the GPU thread's stream
                        records = _records_from_payload(engine, payload_h, H, W, np.asarray(job.boxes),
                                                        [int(v) for v in job.labels], categories, rb, reencode=False)
                free_slots.put(slot)                           # pinned buffers are copied out: recycle the slot
                if on_tile is not None:
                    on_tile(job, label_map, records)
                if save_dir is not None:
                    pool.submit(timed_save, save_dir, job.name, label_map, mapping, records)
            except BaseException as ex:
                errors.append(ex)
                free_slots.put(slot)

    writer_cpu = [0.0]

    def timed_save(*a):
        t = time.time()
        try:
            writers.save_tile(*a)
        except BaseException as ex:
            errors.append(ex)
        with lock:
            writer_cpu[0] += time.time() - t

    loaders = [threading.Thread(target=loader, daemon=True) for _ in range(max(1, loader_threads))]
    pool = ThreadPoolExecutor(max_workers=max(1, writer_threads))
    fin = threading.Thread(target=finisher, args=(pool,), daemon=True)
    for t in loaders:
        t.start()
    fin.start()
    tiles = masks = 0
    t0 = time.time()
    live = len(loaders)
    try:
        while live:
            item = ready.get()
            if item is None:
                live -= 1
                continue
            if errors:
                break
            job, slot, view = item
            H, W = int(view.shape[0]), int(view.shape[1])
            if slot.canvas_d is None or slot.canvas_d.numel() < H * W:
                slot.canvas_d = torch.empty(H * W, dtype=torch.uint8, device=device)
                slot.canvas_h = torch.empty(H * W, dtype=torch.uint8).pin_memory()
            canvas = slot.canvas_d[: H * W].view(H, W)
            boxes = torch.as_tensor(np.asarray(job.boxes))
            labels = torch.as_tensor(np.asarray(job.labels))
            _, payload = tile_outputs(predictor, engine, view.numpy(), boxes, labels, canvas, chunk, instance, capacity_per_mask)
            slot.canvas_h[: H * W].copy_(canvas.view(-1), non_blocking=True)
            slot.event.record(torch.cuda.current_stream(device))
            done_q.put((job, slot, H, W, payload))            # the finisher copies the runs once it knows how many there are
            tiles += 1
            masks += int(boxes.shape[0])
    finally:
        done_q.put(None)
        fin.join()
        t_payload = time.time() - t0                         # every tile's label map and records are on the host
        pool.shutdown(wait=True)
    if errors:
        raise errors[0]
    return {"tiles": tiles, "masks": masks, "seconds": time.time() - t0, "seconds_before_writers_drain": t_payload,
            "writer_cpu_seconds": writer_cpu[0]}

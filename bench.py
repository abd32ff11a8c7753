#!/usr/bin/env python
"""bench.py -- box-prompted masks/sec, SAM ViT-H, 1024x1024 synthetic RS tiles, 32 hboxes per tile.

A "step" is one pass of the hot path over one tile: image encoder + prompt encoder + mask decoder for the
tile's 32 boxes + full-resolution bool masks + the fused semantic label map
(what one iteration of `Generate Dataset/main_sam_hbox_semantic.py:110-216` computes).

  python bench.py [--gpus N] [--steps K] [--warmup W]            our engine (one process per GPU under torchrun)
  python bench.py --impl reference ...                           the reference's CPU path (oracle port) on host cores

`value`   : whole-job masks/s with tiles and boxes already resident in HBM (CUDA events, max over ranks).
`e2e`     : same metric through the drop-in `segment_anything.SamPredictor` with HOST buffers: pinned image and
            boxes copied H2D each step, the driver's 20+12 box chunks, label map copied D2H each step.
`roofline`: the tcgen05 GEMM (dominant kernel): algorithmic FLOPs per launch / mean launch time, measured with
            CUDA events on the launching stream inside the timed region (samrs_profile).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from samrs_b200 import synth  # noqa: E402
from samrs_b200.config import geometry  # noqa: E402
from samrs_b200.stream import broadcast_state_dict, max_over_ranks, shard_indices  # noqa: E402
from samrs_b200.weights import synthetic_state_dict  # noqa: E402

VARIANT = "vit_h"
BOXES = 32
CHUNK = 20                      # the driver's batch_size (main_sam_hbox_semantic.py:91)
METRIC = "box_prompted_masks_per_sec"
N_TILES = 4                     # distinct tiles cycled through (weights alone are 1.3 GB >> 126 MB L2)


def gemm_flops_per_encode(g) -> float:
    """Algorithmic FLOPs (2MNK) of the tensor-core GEMMs of one encode (no padded rows)."""
    D, T = g.embed_dim, 4096
    f = 2.0 * T * 768 * D
    f += g.depth * 2.0 * T * D * (3 * D + D + 4 * D + 4 * D)
    f += 2.0 * T * D * 256 + 2.0 * T * 2304 * 256
    return f


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1590.0))), "measured (sustained bf16 cuBLAS, MEASURED_PEAKS.json)"
    return 1590.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            self._stop.wait(0.05)

    def __enter__(self):
        if self.nv:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._t:
            self._t.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


# ------------------------------------------------------------------------------------------------ reference arm
def cpu_tile(w, g, idx):
    """One tile through the oracle port exactly as the driver issues it: set_image + 20/12-box predict_torch
    chunks + the numpy painter."""
    from oracle import sam_oracle as O
    img = synth.tile(idx)
    boxes = torch.from_numpy(synth.hboxes(idx, BOXES))
    labels = synth.labels(idx, BOXES)
    with torch.no_grad():
        feat = O.set_image(w, g, img)
        seg = np.full((1024, 1024), 255, dtype=np.uint8)
        for s in range(0, BOXES, CHUNK):
            tb = O.apply_boxes(boxes[s:s + CHUNK], (1024, 1024))
            masks, _, _ = O.predict_torch(w, g, feat, None, None, tb, None, False)
            O.painter_reduce(masks[:, 0].numpy(), labels[s:s + CHUNK], seg)
    return seg


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    g = geometry(VARIANT)
    w = synthetic_state_dict(VARIANT, 0)
    budget = float(os.environ.get("SAMRS_REF_BUDGET_S", "240"))
    t_start = time.time()
    warm = min(args.warmup, 1)
    for i in range(warm):
        cpu_tile(w, g, 1000 + i)
    per, done = [], 0
    for i in range(args.steps):
        t0 = time.time()
        cpu_tile(w, g, i)
        per.append(time.time() - t0)
        done += 1
        if time.time() - t_start + per[-1] > budget:
            break
    ms = 1000.0 * float(np.mean(per))
    val = BOXES / (ms / 1000.0)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "masks/s", "n_gpus": args.gpus, "steps": done,
        "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "SAM ViT-H, 1024x1024 synthetic tile, 32 hbox prompts (20+12 chunks), seeded synthetic weights"},
        "cpu_baseline": {"value": val, "unit": "masks/s", "cores": cores, "kind": "port",
                         "sample": f"{done} ViT-H tile(s) x 32 boxes through oracle/sam_oracle.py (torch CPU fp32, {cores} threads)"},
        "e2e": {"value": val, "unit": "masks/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (samrs_b200 has no CPU path; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    import samrs_b200
    from samrs_b200.engine import Engine
    g = geometry(VARIANT)
    # NS tiles in flight per GPU: each has its own engine (activations + weight copy) and CUDA stream, so one tile's
    # kernel tails / launch gaps are filled by the other tile's kernels (tiles are independent)
    NS = max(1, int(os.environ.get("SAMRS_STREAMS", "2")))
    if world > 1:
        sd = broadcast_state_dict(g, lambda: synthetic_state_dict(VARIANT, 0), device, src=0)
    else:
        sd = synthetic_state_dict(VARIANT, 0)
    engines = []
    for _ in range(NS):
        e_ = Engine(VARIANT, device)
        e_.load_state_dict(sd)
        engines.append(e_)
    eng = engines[0]
    del sd
    torch.cuda.empty_cache()
    streams = [torch.cuda.Stream(device=device) for _ in range(NS)]

    # the drop-in predictors share the engines (same weights, no further copies)
    sys.path.insert(0, samrs_b200.DROPIN_PATH)
    from segment_anything import SamPredictor
    from segment_anything.modeling import Sam
    predictors = []
    for e_ in engines:
        sam = Sam(g)
        sam.engine, sam._device = e_, e_.device
        predictors.append(SamPredictor(sam))

    # this rank's tiles: rank r takes tiles r, r+world, ... (files[rank::world])
    idxs = shard_indices(N_TILES * world, rank, world)
    tiles_h = [torch.from_numpy(synth.tile(i)).pin_memory() for i in idxs]
    boxes_h = [torch.from_numpy(synth.hboxes(i, BOXES)).pin_memory() for i in idxs]
    labels_h = [torch.from_numpy(synth.labels(i, BOXES)).to(torch.int32).pin_memory() for i in idxs]
    tiles_d = [t.to(device) for t in tiles_h]
    boxes_d = [b.to(device) for b in boxes_h]
    labels_d = [l.to(device) for l in labels_h]
    canvases = [torch.empty((1024, 1024), dtype=torch.uint8, device=device) for _ in range(NS)]

    VARIANT_STEP = os.environ.get("SAMRS_BENCH_STEP", "")     # diagnostics only; the reported line uses the default

    def step_resident(i):
        nonlocal NS
        j, k = i % N_TILES, i % NS
        en, canvas = engines[k], canvases[k]
        with torch.cuda.stream(streams[k]):
            en.encode(tiles_d[j])
            if VARIANT_STEP == "chunks":             # diagnostic: the driver's 20 + 12 chunking on resident inputs
                canvas.fill_(255)
                for s in range(0, BOXES, CHUNK):
                    low, _ = en.decode(boxes=boxes_d[j][s:s + CHUNK], multimask_output=False)
                    en.postprocess(low, (1024, 1024), (1024, 1024))
                    en.semantic_reduce(low, labels_d[j][s:s + CHUNK], canvas)
                return
            low, _ = en.decode(boxes=boxes_d[j], multimask_output=False)
            en.postprocess(low, (1024, 1024), (1024, 1024))
            canvas.fill_(255)
            en.semantic_reduce(low, labels_d[j], canvas)

    # e2e output ring: the host consumes label maps with a lag of RING steps, so the only host wait in a step is for
    # the D2H of the tile RING steps back (device buffers are reused in stream order and need no host sync)
    RING = 4
    outs_h = [torch.empty((1024, 1024), dtype=torch.uint8).pin_memory() for _ in range(RING)]
    out_done = [None] * RING

    def step_e2e(i):
        j, k, r = i % N_TILES, i % NS, i % RING
        en, canvas, predictor = engines[k], canvases[k], predictors[k]
        if out_done[r] is not None:
            out_done[r].synchronize()                                    # the host consumed this slot's previous label map
        with torch.cuda.stream(streams[k]):
            img = tiles_h[j].numpy()                                    # host HWC uint8 (pinned)
            predictor.set_image(img)                                     # H2D 3 MiB + encoder
            bx = boxes_h[j].to(device, non_blocking=True)                # H2D 512 B
            lb = labels_h[j].to(device, non_blocking=True)
            canvas.fill_(255)
            for s in range(0, BOXES, CHUNK):                             # the driver's 20 + 12 chunks
                tb = predictor.transform.apply_boxes_torch(bx[s:s + CHUNK], img.shape[:2])
                _, _, low = predictor.predict_torch(None, None, boxes=tb, mask_input=None, multimask_output=False)
                en.semantic_reduce(low, lb[s:s + CHUNK], canvas)
            outs_h[r].copy_(canvas, non_blocking=True)                   # D2H 1 MiB label map
            ev = torch.cuda.Event()
            ev.record(streams[k])
            out_done[r] = ev

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, profile=False):
        barrier()
        if profile:
            for e_ in engines:
                e_.profile_begin()
        l0 = sum(e_.launch_count() for e_ in engines)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        cur = torch.cuda.current_stream()
        a.record(cur)
        for st_ in streams:
            st_.wait_stream(cur)
        for i in range(steps):
            fn(i)
        for st_ in streams:
            cur.wait_stream(st_)
        b.record(cur)
        torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        prof = None
        if profile:
            prof = {}
            for e_ in engines:
                for k_, (ms_, n_) in e_.profile_end().items():
                    o = prof.get(k_, (0.0, 0))
                    prof[k_] = (o[0] + ms_, o[1] + n_)
        launches = sum(e_.launch_count() for e_ in engines) - l0
        if world > 1:
            ms = max_over_ranks(ms, device)
        barrier()
        return ms, launches, prof

    for i in range(max(args.warmup, 3)):
        step_resident(i)
        step_e2e(i)
    with ClockSampler(local) as clk:
        ms_res, launches, _ = timed(step_resident, args.steps)
        ms_e2e, _, _ = timed(step_e2e, args.steps)
    clocks = clk.summary()
    # roofline pass: the same step with ONE tile in flight, so that each kernel's CUDA-event duration is its own
    # (with several tiles in flight kernels of different tiles share the SMs and per-kernel times overlap)
    ns_saved, NS = NS, 1
    prof_steps = max(2, min(args.steps, 6))
    ms_single, _, prof = timed(step_resident, prof_steps, profile=True)
    NS = ns_saved

    value = world * BOXES * args.steps / (ms_res / 1000.0)
    e2e = world * BOXES * args.steps / (ms_e2e / 1000.0)
    peak, peak_src = measured_peaks()
    gemm_ms, gemm_n = prof["gemm_tc"]
    flops_per_launch = gemm_flops_per_encode(g) * prof_steps / max(gemm_n, 1)
    achieved = flops_per_launch / (gemm_ms / max(gemm_n, 1) / 1000.0) / 1e12 if gemm_ms > 0 else 0.0
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "gemm_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get("dram_bytes_per_launch")
    step_ms = ms_res / args.steps
    single_ms = ms_single / prof_steps
    shares = {k: round(v[0] / prof_steps, 4) for k, v in prof.items()}

    line = {
        "metric": METRIC, "value": value, "unit": "masks/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f16 tensor-core operands, f32 accumulate / residual / softmax / decoder", "data": "synthetic",
        "config": {"workload": "SAM ViT-H, 1024x1024 synthetic RS tile, 32 hbox prompts per tile (BASELINE.json configs[1])",
                   "tiles_per_rank_cycled": N_TILES, "tiles_in_flight_per_gpu": NS, "l2": "working set (1.3 GB fp16 weights) exceeds the 126 MB L2",
                   "parallelism": f"tile-sharded dp{world}, weights broadcast once over NCCL" if world > 1 else "single GPU",
                   "step": "encode + decode(32) + bool masks + fused label map"},
        "gpu_launches": launches,
        "e2e": {"value": e2e, "unit": "masks/s", "h2d_bytes_per_step": 1024 * 1024 * 3 + BOXES * 16 + BOXES * 4,
                "d2h_bytes_per_step": 1024 * 1024, "ms_per_step": ms_e2e / args.steps,
                "path": "segment_anything.SamPredictor.set_image + predict_torch (20+12 chunks) + semantic_reduce"},
        "roofline": {"bound": "tensor", "kernel": "gemm_tc2_kernel / gemm_tc_kernel (tcgen05 cta_group::2 / ::1, fp16 in, fp32 acc): every encoder GEMM launch", "achieved": achieved, "peak": peak,
                     "unit": "TFLOP/s", "frac": achieved / peak, "peak_source": peak_src, "traffic": traffic,
                     "launches_per_step": gemm_n / prof_steps, "share_of_step": (gemm_ms / prof_steps) / single_ms,
                     "measured": "CUDA events around every launch on its stream, one tile in flight"},
        "single_tile_in_flight": {"ms_per_step": single_ms, "ms_per_step_by_kernel": shares},
        "clocks": clocks,
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        w = synthetic_state_dict(VARIANT, 0)
        t0 = time.time()
        cpu_tile(w, g, 0)
        dt = time.time() - t0
        line["cpu_baseline"] = {"value": BOXES / dt, "unit": "masks/s", "cores": cores, "kind": "port",
                                "sample": f"1 ViT-H tile x 32 boxes through oracle/sam_oracle.py (torch CPU fp32, {cores} threads, {dt:.1f} s)"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

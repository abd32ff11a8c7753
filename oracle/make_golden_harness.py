"""Fixtures for the run-unchanged harness: the three drivers executed over the REFERENCE package on the CPU.

Build container only:   python oracle/make_golden_harness.py [stage_dir]

`samrs_b200.harness` stages the synthetic DIOR / FAIR1M / HRSC trees and the seeded ViT-H checkpoint, then runs the
UNMODIFIED driver scripts with `sys.path = [/root/reference/Generate Dataset, ...]`, i.e. over the reference's own
`segment_anything`.  The drivers say `.cuda()` / `device="cuda"` and this container has no GPU, so for THIS run only
`Tensor.cuda` / `Module.to("cuda")` are mapped to the CPU (the arithmetic is the reference's fp32 CPU path, 8 threads).
What is committed (tests/golden/harness/): the gray label PNGs, per-instance areas of the pickles, and the rbox driver's
`sam_ins_rbox.json`; the GPU test runs the same drivers over the drop-in on a B200 and bounds the differing pixels.
"""
from __future__ import annotations

import json
import os
import pickle
import shutil
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GD = "/root/reference/Generate Dataset"

from samrs_b200 import harness  # noqa: E402


def cpu_shim():
    torch.Tensor.cuda = lambda self, *a, **k: self
    orig_to = torch.nn.Module.to

    def to(self, *a, **k):
        a = tuple("cpu" if (isinstance(x, str) and x.startswith("cuda")) else x for x in a)
        if isinstance(k.get("device"), str) and k["device"].startswith("cuda"):
            k["device"] = "cpu"
        return orig_to(self, *a, **k)
    torch.nn.Module.to = to


def main():
    stage = os.path.abspath(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, ".stage_tmp", "ref"))
    out = os.path.join(ROOT, "tests", "golden", "harness")
    os.makedirs(out, exist_ok=True)
    torch.set_num_threads(8)
    cpu_shim()
    redirect = harness.stage_all(stage)
    ds = redirect["/root/dataset"]
    meta = {"torch": torch.__version__, "threads": 8, "drivers": {}}
    for script, argv in (("main_sam_hbox_semantic.py", []), ("main_sam_rhbox_semantic.py", []),
                         ("main_sam_rbox_mask_instance.py", ["--show", "False"])):
        t0 = time.time()
        harness.run_driver(os.path.join(GD, script), argv, package_dir=GD, redirect=redirect)
        meta["drivers"][script] = {"argv": argv, "seconds": round(time.time() - t0, 1)}
        print(f"{script}: {time.time() - t0:.1f} s", flush=True)
    for tag, save in (("hbox", os.path.join(ds, "dior", "hbox_segs_test_init")),
                      ("rhbox", os.path.join(ds, "fair1m_1024", "trainval", "rhbox_segs_init"))):
        areas = {}
        for f in sorted(os.listdir(os.path.join(save, "gray"))):
            shutil.copyfile(os.path.join(save, "gray", f), os.path.join(out, f"{tag}_gray_{f}"))
            with open(os.path.join(save, "ins", f[:-4] + ".pkl"), "rb") as fh:
                recs = pickle.load(fh)
            areas[f[:-4]] = {"size": [int(r["size"]) for r in recs], "label": [int(r["label"]) for r in recs],
                             "category": [r["category"] for r in recs], "keys": sorted(recs[0].keys())}
        meta[tag] = areas
    shutil.copyfile(os.path.join(redirect["/root/dw"], "samrs", "work_dir", "hrsc", "json", "sam_ins_rbox.json"),
                    os.path.join(out, "rbox_sam_ins_rbox.json"))
    with open(os.path.join(out, "meta.json"), "w") as f:
        json.dump(meta, f, indent=1)
    print("fixtures ->", out, sorted(os.listdir(out)))


if __name__ == "__main__":
    main()

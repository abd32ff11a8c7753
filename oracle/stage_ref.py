"""Stage the parts of the reference that tests and the reference arm EXECUTE into oracle/_ref/ (git-ignored).

Build container only (`/root/reference` does not exist on the GPU box; `oracle/_ref/` is git-ignored but not
gpurun-ignored, so the staged files travel with the snapshot exactly like our own built `.so`).  Nothing staged here is
ever imported by the product (`samrs_b200/`): the staged package is the checker / the CPU baseline, the staged driver
scripts are what `samrs_b200.harness` runs UNMODIFIED against the drop-in package.

    oracle/_ref/GD/segment_anything/      the reference's vendored package  -> `bench.py --impl reference`, goldens
    oracle/_ref/GD/main_sam_*.py          the three generation drivers      -> tests/test_harness_gpu.py
    oracle/_ref/GD/{loaddata,mapping,instance_to_json}.py, utils/           their local imports

Called by `__graft_entry__.build()` when /root/reference is present.
"""
from __future__ import annotations

import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference/Generate Dataset"
DST = os.path.join(ROOT, "oracle", "_ref", "GD")

FILES = ["main_sam_hbox_semantic.py", "main_sam_rhbox_semantic.py", "main_sam_rbox_mask_instance.py",
         "loaddata.py", "mapping.py", "instance_to_json.py"]
DIRS = ["segment_anything", "utils"]


def staged() -> bool:
    return all(os.path.exists(os.path.join(DST, f)) for f in FILES + DIRS)


def stage() -> bool:
    """Copy the files (byte-identical) if the reference tree is present; returns whether oracle/_ref/GD is complete."""
    if not os.path.isdir(SRC):
        return staged()
    os.makedirs(DST, exist_ok=True)
    for f in FILES:
        shutil.copyfile(os.path.join(SRC, f), os.path.join(DST, f))
    for d in DIRS:
        dst = os.path.join(DST, d)
        if os.path.isdir(dst):
            shutil.rmtree(dst)
        shutil.copytree(os.path.join(SRC, d), dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    return staged()


if __name__ == "__main__":
    print("staged" if stage() else "reference tree absent and nothing staged", DST)

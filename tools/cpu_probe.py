"""Probe the GPU box's host for the reference arm: usable cores and how the reference's ViT-H set_image scales with threads."""
import os, sys, time, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref", "GD"))
out = {"cpu_count": os.cpu_count(), "affinity": len(os.sched_getaffinity(0))}
for p in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
    if os.path.exists(p):
        out[p] = open(p).read().strip()
try:
    out["lscpu"] = [l for l in os.popen("lscpu").read().splitlines() if any(k in l for k in ("Model name", "Socket", "Core(s)", "Thread(s)", "NUMA node(s)"))]
except Exception:
    pass
from segment_anything import SamPredictor
from segment_anything.build_sam import _build_sam
from samrs_b200 import synth
from samrs_b200.weights import synthetic_state_dict
sam = _build_sam(1280, 32, 16, [7, 15, 23, 31])
sam.load_state_dict(synthetic_state_dict("vit_h", 0))
pred = SamPredictor(sam)
img = synth.tile(0)
for th in [int(a) for a in sys.argv[1:]] or [32, 64]:
    torch.set_num_threads(th)
    with torch.no_grad():
        t0 = time.time(); pred.set_image(img); t1 = time.time() - t0
        t0 = time.time(); pred.set_image(img); t2 = time.time() - t0
    out[f"set_image_s_threads_{th}"] = [round(t1, 2), round(t2, 2)]
    print(th, t1, t2, flush=True)
print(json.dumps(out, indent=1))
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "cpu_probe.json"), "w"), indent=1)

#!/bin/bash
timeout 300 python -m pytest tests/test_resize.py tests/test_abi.py -q -m gpu 2>&1 | tail -4
timeout 100 python - <<'PY'
import time, numpy as np, torch
from PIL import Image
from samrs_b200.engine import Engine
eng = Engine("vit_t64", "cuda:0")
img = np.random.default_rng(0).integers(0, 256, (800, 800, 3), dtype=np.uint8)
d = torch.from_numpy(img).cuda()
for _ in range(3): eng.resize_image(d, (1024, 1024))
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(50): eng.resize_image(d, (1024, 1024))
b.record(); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10): np.array(Image.fromarray(img).resize((1024, 1024), Image.BILINEAR))
t1 = time.perf_counter()
print(f"800x800 -> 1024x1024: device {a.elapsed_time(b) / 50 * 1e3:.1f} us, PIL on the host {(t1 - t0) / 10 * 1e3:.2f} ms")
PY

"""Times the on-device instance payload (RLE + area) of one tile's 32 masks against the reference route
(D2H of bool masks + per-mask numpy encode, oracle/rle_oracle.py standing in for pycocotools)."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from samrs_b200.engine import Engine
from samrs_b200 import synth
from oracle import rle_oracle

eng = Engine("vit_t64", "cuda:0")
B = 32
# box-like logits: positive inside a synthetic box, negative outside, plus noise -> realistic run counts
low = torch.full((B, 1, 256, 256), -4.0)
for b, (x0, y0, x1, y1) in enumerate(synth.hboxes(0, B)):
    low[b, 0, int(y0) // 4:int(y1) // 4 + 1, int(x0) // 4:int(x1) // 4 + 1] = 4.0
low += torch.randn(low.shape, generator=torch.Generator().manual_seed(0))
low = low.cuda()
masks = eng.postprocess(low, (1024, 1024), (1024, 1024))

def timed(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3

us_fused = timed(lambda: eng.rle_encode(low_res=low, capacity=1 << 22))
us_masks = timed(lambda: eng.rle_encode(masks=masks, capacity=1 << 22))
us_post = timed(lambda: eng.postprocess(low, (1024, 1024), (1024, 1024)))
counts, offsets, area = eng.rle_encode(low_res=low, capacity=1 << 22)
torch.cuda.synchronize()
runs = int(offsets[-1])
t0 = time.perf_counter()
m_np = masks.reshape(B, 1024, 1024).cpu().numpy()
t1 = time.perf_counter()
c, o, a = rle_oracle.encode_batch(m_np)
t2 = time.perf_counter()
assert np.array_equal(o, offsets.cpu().numpy()) and np.array_equal(c, counts[:runs].cpu().numpy())
t3 = time.perf_counter()
payload = counts[:runs].cpu(), offsets.cpu(), area.cpu()
t4 = time.perf_counter()
print(f"masks {B}, runs {runs}, mean area {float(area.float().mean()):.0f}")
print(f"device fused (low-res -> runs): {us_fused:8.1f} us   [{B * 256 * 256 * 4 / us_fused / 1e3:.1f} GB/s of logits, {B * 1024 * 1024 / 8 / us_fused / 1e3:.1f} GB/s of packed bits]")
print(f"device from bool masks:         {us_masks:8.1f} us   [{B * 1024 * 1024 / us_masks / 1e3:.1f} GB/s of mask bytes]   (+ postprocess {us_post:.1f} us to make them)")
print(f"reference route: D2H of {B} MiB masks {1e3 * (t1 - t0):.1f} ms + numpy encode {1e3 * (t2 - t1):.1f} ms;  D2H of the device payload {1e3 * (t4 - t3):.2f} ms ({(runs * 4 + B * 16) / 1024:.1f} KiB)")

"""Times the tcgen05 GEMM variants (1-CTA / CTA-pair x N tile) on the ViT-H shapes with CUDA events."""
import math, sys, os, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from samrs_b200.engine import Engine

eng = Engine("vit_t64", "cuda:0")
shapes = {"qkv": (4096, 3840, 1280, True, False), "proj": (4096, 1280, 1280, False, False),
          "lin1": (4096, 5120, 1280, True, True), "lin1_nogelu": (4096, 5120, 1280, True, False), "lin2": (4096, 1280, 5120, False, False)}
cfgs = [int(c) for c in sys.argv[1:]] or [128, 160, 224, 256, 1160, 1224, 1256]
res = {}
for name, (M, N, K, half, gelu) in shapes.items():
    A = (torch.randn(M, K, device="cuda") * 1.0).half()
    B = (torch.randn(N, K, device="cuda") / math.sqrt(K)).half()
    bias = torch.randn(N, device="cuda")
    r = None if half else torch.randn(M, N, device="cuda")
    # library yardstick (not a product path): cuBLAS on the same shape, no bias / activation / residual
    Bt = B.t()
    for _ in range(3): torch.matmul(A, Bt)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20): torch.matmul(A, Bt)
    b.record(); torch.cuda.synchronize()
    us = a.elapsed_time(b) / 20 * 1000
    res[f"{name}/cublas"] = (round(us, 1), round(2.0 * M * N * K / us / 1e6, 1))
    print(f"{name:11s} cuBLAS   : {us:7.1f} us  {2.0*M*N*K/us/1e6:7.1f} TFLOP/s", flush=True)
    for c in cfgs:
        if half and (c % 1000 % 32 != 0 or 3000 <= c < 4000):
            continue                        # fp32-epilogue-only variants: tiles not a multiple of 32 wide, stream-K (3000 + bn)
        if 3000 <= c < 4000:                # stream-K exists for the in-place reduce-add epilogue only
            x = torch.randn(M, N, device="cuda")
            for _ in range(3): eng.test_gemm(A, B, out_half=False, bias=bias, gelu=gelu, force_bn=c, res=x, out=x)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(20): eng.test_gemm(A, B, out_half=False, bias=bias, gelu=gelu, force_bn=c, res=x, out=x)
            b.record(); torch.cuda.synchronize()
            us = a.elapsed_time(b) / 20 * 1000
            res[f"{name}/{c}"] = (round(us, 1), round(2.0 * M * N * K / us / 1e6, 1))
            print(f"{name:11s} cfg {c:5d} [inplace]: {us:7.1f} us  {2.0*M*N*K/us/1e6:7.1f} TFLOP/s", flush=True)
            continue
        for _ in range(3):
            eng.test_gemm(A, B, out_half=half, bias=bias, res=r, gelu=gelu, force_bn=c)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20):
            eng.test_gemm(A, B, out_half=half, bias=bias, res=r, gelu=gelu, force_bn=c)
        b.record(); torch.cuda.synchronize()
        us = a.elapsed_time(b) / 20 * 1000
        res[f"{name}/{c}"] = (round(us, 1), round(2.0 * M * N * K / us / 1e6, 1))
        print(f"{name:11s} cfg {c:5d}: {us:7.1f} us  {2.0*M*N*K/us/1e6:7.1f} TFLOP/s", flush=True)
        if not half:      # epilogue flavours of the fp32-output GEMMs: plain store / residual read (above) / in-place TMA reduce-add
            for label, kw in (("store", dict(res=None)), ("inplace", dict(res=r, out=r))):
                for _ in range(3): eng.test_gemm(A, B, out_half=half, bias=bias, gelu=gelu, force_bn=c, **kw)
                torch.cuda.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(20): eng.test_gemm(A, B, out_half=half, bias=bias, gelu=gelu, force_bn=c, **kw)
                b.record(); torch.cuda.synchronize()
                us2 = a.elapsed_time(b) / 20 * 1000
                print(f"{name:11s} cfg {c:5d} [{label:7s}]: {us2:7.1f} us", flush=True)
json.dump(res, open("gpurun_out/gemm_sweep.json", "w"))

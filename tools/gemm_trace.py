"""Pipeline trace of the CTA-pair GEMM (CTA 0): clock64 at prologue / per-k-block full-barrier / tile commit / epilogue."""
import ctypes, math, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from samrs_b200.engine import Engine, load_library
eng = Engine("vit_t64", "cuda:0")
lib = load_library()
lib.samrs_test_set_gemm_trace.argtypes = [ctypes.c_void_p]
buf = torch.zeros(4096, dtype=torch.int64, device="cuda")
for name, (M, N, K, half, useres) in {"qkv": (4096, 3840, 1280, True, False), "lin2": (4096, 1280, 5120, False, True)}.items():
    A = torch.randn(M, K, device="cuda").half(); B = (torch.randn(N, K, device="cuda") / math.sqrt(K)).half()
    bias = torch.randn(N, device="cuda"); r = torch.randn(M, N, device="cuda") if useres else None
    for cfg in (1256,):
        for _ in range(3): eng.test_gemm(A, B, out_half=half, bias=bias, res=r, force_bn=cfg)
        torch.cuda.synchronize(); buf.zero_()
        lib.samrs_test_set_gemm_trace(buf.data_ptr())
        eng.test_gemm(A, B, out_half=half, bias=bias, res=r, force_bn=cfg)
        torch.cuda.synchronize()
        lib.samrs_test_set_gemm_trace(None)
        t = buf.cpu().tolist(); t0 = t[0]
        print(f"== {name} cfg {cfg}: start->setup done {t[1]-t0}, loops done {t[2]-t0}, after final cluster sync {t[3]-t0}")
        for ti in range(6):
            base = 16 + ti * 64
            if t[base] == 0: break
            kbs = [t[base + 2 + k] - t0 for k in range(40) if t[base + 2 + k]]
            d = [kbs[i + 1] - kbs[i] for i in range(len(kbs) - 1)]
            print(f" tile {ti}: mma_start {t[base]-t0} tempty_ok {t[base+1]-t0} first_full {kbs[0] if kbs else None} "
                  f"kb deltas(first 40) min {min(d) if d else 0} med {sorted(d)[len(d)//2] if d else 0} max {max(d) if d else 0} "
                  f"issue_done {t[base+60]-t0} | epi start {t[base+61]-t0 if t[base+61] else None} epi end {t[base+62]-t0 if t[base+62] else None}")
            if ti == 1: print("   kb deltas:", d)

"""Pipeline experiments on the production 1-CTA GEMM: clock64 trace of CTA 0 plus timing with parts of the pipeline removed
(dbg_mode bit0 = no TMA loads, bit1 = no MMAs, bit2 = no epilogue) to see which stage bounds the k-block period."""
import ctypes, math, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from samrs_b200.engine import Engine, load_library
eng = Engine("vit_t64", "cuda:0")
lib = load_library()
lib.samrs_test_set_gemm_trace.argtypes = [ctypes.c_void_p]
lib.samrs_test_set_gemm_mode.argtypes = [ctypes.c_int]
buf = torch.zeros(4096, dtype=torch.int64, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
shapes = {"qkv": (4096, 3840, 1280, True, False), "proj": (4096, 1280, 1280, False, True), "lin1": (4096, 5120, 1280, True, False),
          "lin1_gelu": (4096, 5120, 1280, True, False),
          "lin2": (4096, 1280, 5120, False, True)}
cfgs = {"qkv": (1224,), "proj": (1160,), "lin1": (1224,), "lin1_gelu": (1224,), "lin2": (1160,)}
for name, (M, N, K, half, useres) in shapes.items():
    A = torch.randn(M, K, device="cuda").half(); B = (torch.randn(N, K, device="cuda") / math.sqrt(K)).half()
    bias = torch.randn(N, device="cuda"); r = torch.randn(M, N, device="cuda") if useres else None
    gelu = name.endswith("gelu")
    for cfg in cfgs[name]:
        line = f"{name:5s} cfg {cfg}:"
        for mode in ((0, 1, 2, 4, 5, 6, 3) if cfg < 1000 else (0,)):
            lib.samrs_test_set_gemm_mode(mode)
            for _ in range(3): eng.test_gemm(A, B, out_half=half, bias=bias, res=r, out=r, gelu=gelu, force_bn=cfg)
            ts = []
            for _ in range(10):
                flush.zero_()
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                e0.record(); eng.test_gemm(A, B, out_half=half, bias=bias, res=r, out=r, gelu=gelu, force_bn=cfg); e1.record()
                torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
            line += f"  m{mode} {sorted(ts)[len(ts)//2]:6.1f}us"
        lib.samrs_test_set_gemm_mode(0)
        print(line, "  (m0 full, m1 noTMA, m2 noMMA, m4 noEpi, m5 noTMA+noEpi, m6 noMMA+noEpi, m3 noTMA+noMMA)")
        if True:
            torch.cuda.synchronize(); buf.zero_(); buf[10] = 1 << 62
            if os.environ.get("TRACE_HOT"):          # trace the launch that follows 40 back-to-back ones (sustained clocks)
                for _ in range(40): eng.test_gemm(A, B, out_half=half, bias=bias, res=r, out=r, gelu=gelu, force_bn=cfg)
            lib.samrs_test_set_gemm_trace(buf.data_ptr())
            eng.test_gemm(A, B, out_half=half, bias=bias, res=r, out=r, gelu=gelu, force_bn=cfg)
            torch.cuda.synchronize()
            lib.samrs_test_set_gemm_trace(None)
            t = buf.cpu().tolist(); t0 = t[1] - 1400 if t[0] == 0 else t[0]
            print(f"   trace: setup done {t[1]-t0}, loops done {t[2]-t0}, end {t[3]-t0} clk | wall: grid {t[11]-t[10]} ns, CTA0 {t[13]-t[12]} ns "
                  f"(starts {t[12]-t[10]} ns after the first CTA) -> {(t[3]-t0)/max(1,t[13]-t[12]):.2f} GHz")
            for ti in range(6):
                base = 16 + ti * 64
                if t[base] == 0: break
                kbs = [t[base + 2 + k] - t0 for k in range(40) if t[base + 2 + k]]
                d = [kbs[i + 1] - kbs[i] for i in range(len(kbs) - 1)]
                print(f"   tile {ti}: mma_start {t[base]-t0} first_full {kbs[0] if kbs else None} kb min {min(d) if d else 0} med {sorted(d)[len(d)//2] if d else 0} "
                      f"max {max(d) if d else 0} issue_done {t[base+60]-t0} | epi {t[base+61]-t0 if t[base+61] else None} -> {t[base+62]-t0 if t[base+62] else None}")
            if not half:
                for ci in range(6):
                    e = t[2048 + ci * 8: 2048 + ci * 8 + 7]
                    if e[0] == 0: break
                    print(f"   epi chunk {ci}: start {e[0]-t0}  ld {e[1]-e[0]}  math {e[2]-e[1]}  wait_read {e[3]-e[2]}  sts {e[4]-e[3]}  fence {e[5]-e[4]}  tma_issue {e[6]-e[5]}")

import math, sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from samrs_b200.engine import Engine
eng = Engine("vit_t64", "cuda:0")
def run(M,N,K,half,res,bias,cfg,iters=20):
    A = torch.randn(M, K, device="cuda").half(); B = (torch.randn(N, K, device="cuda")/math.sqrt(K)).half()
    bi = torch.randn(N, device="cuda") if bias else None
    r = torch.randn(M, N, device="cuda") if res else None
    for _ in range(3): eng.test_gemm(A,B,out_half=half,bias=bi,res=r,force_bn=cfg)
    torch.cuda.synchronize()
    a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): eng.test_gemm(A,B,out_half=half,bias=bi,res=r,force_bn=cfg)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b)/iters*1000
for cfg in (160, 1256):
    for (M,N) in ((4096,1280),(4096,3840),(8192,1280)):
        for K in (64, 320, 1280, 5120):
            row=[]
            for (half,res,bias) in ((True,False,False),(False,False,False),(False,True,True)):
                row.append(run(M,N,K,half,res,bias,cfg))
            print(f"cfg {cfg} M{M} N{N} K{K}: half {row[0]:6.1f}  f32 {row[1]:6.1f}  f32+res+bias {row[2]:6.1f} us", flush=True)

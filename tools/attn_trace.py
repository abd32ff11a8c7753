"""Pipeline trace (clock64, CTA 0) of the attention kernel on a ViT-H-shaped qkv activation."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from samrs_b200.config import SamGeometry
from samrs_b200.engine import Engine, load_library
g = SamGeometry("h_like", 1280, 1, 16, (0,))
eng = Engine(g, "cuda:0")
lib = load_library()
lib.samrs_test_set_attn_trace.argtypes = [ctypes.c_void_p]
qkv = torch.randn(4096, 3840, device="cuda").half()
for glob in (False, True):
    S = 64 if glob else 14
    rph = torch.randn(2 * S - 1, 80, device="cuda") * 0.1; rpw = torch.randn(2 * S - 1, 80, device="cuda") * 0.1
    for _ in range(2): eng.test_attention(qkv, rph, rpw, glob)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): eng.test_attention(qkv, rph, rpw, glob)
    b.record(); torch.cuda.synchronize()
    print(f"== global={glob}: {a.elapsed_time(b) / 10 * 1000:.1f} us per call (incl. rel-pos GEMM)")
    buf = torch.zeros(4096, dtype=torch.int64, device="cuda")
    lib.samrs_test_set_attn_trace(buf.data_ptr())
    eng.test_attention(qkv, rph, rpw, glob)
    torch.cuda.synchronize()
    lib.samrs_test_set_attn_trace(None)
    t = buf.cpu().tolist()
    t0 = min(x for x in t if x > 0)
    names = {0: "prod:start", 1: "prod:q_empty ok", 2: "prod:k_empty ok", 3: "prod:v_empty ok", 8: "mma:start", 9: "mma:q_full", 10: "mma:k_full",
             11: "mma:S0 issued", 12: "mma:S1 issued", 13: "mma:v_full", 14: "mma:p_full0", 15: "mma:p_full1",
             20: "wg0:start", 21: "wg0:rel loaded", 22: "wg0:s_full", 23: "wg0:pass1 done", 24: "wg0:P stored", 25: "wg0:o_full", 26: "wg0:out stored",
             30: "wg1:start", 31: "wg1:rel loaded", 32: "wg1:s_full", 33: "wg1:pass1 done", 34: "wg1:P stored", 35: "wg1:o_full", 36: "wg1:out stored"}
    for ui in range(3):
        ev = sorted((t[64 * ui + k] - t0, n) for k, n in names.items() if t[64 * ui + k])
        print(f" unit {ui}: " + "  ".join(f"{n}@{c}" for c, n in ev))

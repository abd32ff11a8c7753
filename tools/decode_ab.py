"""A/B timing of the mask decoder alone (32 boxes, ViT-H engine, random features): CUDA events over 20 graph replays."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from samrs_b200 import synth
from samrs_b200.engine import Engine
from samrs_b200.weights import synthetic_state_dict
eng = Engine("vit_h", "cuda:0")
eng.load_state_dict(synthetic_state_dict("vit_h", 0))
g = torch.Generator(device="cuda").manual_seed(0)
eng.set_features(torch.randn((1, 256, 64, 64), device="cuda", generator=g))
for B in (32, 20, 12, 64):
    boxes = torch.from_numpy(synth.hboxes(0, B)).cuda()
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        for _ in range(4):
            eng.decode(boxes=boxes, multimask_output=False)
        st.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(st)
        for _ in range(20):
            eng.decode(boxes=boxes, multimask_output=False)
        b.record(st)
        st.synchronize()
    print(f"{os.environ.get('SAMRS_LIB', 'libsamrs_b200.so')}: decode B={B}: {a.elapsed_time(b) / 20 * 1000:.1f} us")

"""One ViT-H tile (encode + 32-box decode + epilogue) bracketed by cudaProfilerStart/Stop, for ncu:
   ncu --profile-from-start off ... python tools/profile_step.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from samrs_b200 import synth  # noqa: E402
from samrs_b200.engine import Engine  # noqa: E402
from samrs_b200.weights import synthetic_state_dict  # noqa: E402

variant = sys.argv[1] if len(sys.argv) > 1 else "vit_h"
eng = Engine(variant, "cuda:0")
eng.load_state_dict(synthetic_state_dict(variant, 0))
eng.set_graphs(False)                 # ncu lists plain launches; the graph replays the same kernels
img = torch.from_numpy(synth.tile(0)).cuda()
boxes = torch.from_numpy(synth.hboxes(0, 32)).cuda()
labels = torch.from_numpy(synth.labels(0, 32)).to(torch.int32).cuda()
canvas = torch.full((1024, 1024), 255, dtype=torch.uint8, device="cuda")


def step():
    eng.encode(img)
    low, _ = eng.decode(boxes=boxes, multimask_output=False)
    eng.postprocess(low, (1024, 1024), (1024, 1024))
    eng.semantic_reduce(low, labels, canvas)
    counts, offsets, _ = eng.rle_encode(low_res=low, capacity=1 << 23)
    eng.rle_strings(counts, offsets)


step()
step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()

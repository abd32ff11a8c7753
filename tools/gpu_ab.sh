#!/bin/bash
mkdir -p gpurun_out
echo "--- merged PV"; timeout 300 python -m pytest tests/test_gpu_kernels.py -q -x -k "attention" 2>&1 | tail -2
timeout 100 python tools/attn_trace.py 2>&1 | grep "=="
echo "--- split PV"; SAMRS_ATTN_PV_SPLIT=1 timeout 100 python tools/attn_trace.py 2>&1 | grep "=="
timeout 300 python -m pytest tests/test_rle.py -q -m gpu 2>&1 | tail -2

#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_kernels.py -q -x -k "not gemm_" 2>&1 | tail -2
bash tools/gpu_ab3.sh "$@"

#!/bin/bash
# cluster-of-4 multicast GEMM: bounded correctness tests, then the sweep against the pair kernel and cuBLAS
mkdir -p gpurun_out
SAMRS_VERBOSE=1 timeout 200 python -m pytest tests/test_gpu_kernels.py -q -x -k "cluster_of_four" 2>&1 | tail -8
SAMRS_VERBOSE=1 timeout 300 python tools/gemm_sweep.py 1224 4224 1160 4160 4256 2>&1 | grep -v "store\|inplace" > gpurun_out/gemm_sweep_quad.txt; cat gpurun_out/gemm_sweep_quad.txt

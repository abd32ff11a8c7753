#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -x -k "gemm" 2>&1 | tail -8
python tools/gemm_sweep.py 2>&1 | tail -30
python tools/gemm_trace.py 2>&1 | tail -14
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x 2>&1 | tail -5
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json

#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -x -k "gemm" 2>&1 | tail -2
timeout 200 python tools/gemm_sweep.py 1160 1224 1256 2>&1 | grep -E "qkv|lin1"
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x 2>&1 | tail -2
run() { timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('$1', 'value', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'single', round(d['single_tile_in_flight']['ms_per_step'],3), 'gemm', d['single_tile_in_flight']['ms_per_step_by_kernel']['gemm_tc'], 'roof', round(d['roofline']['achieved'],1))" || tail -3 gpurun_out/bench.err; }
run "12 epilogue warps"

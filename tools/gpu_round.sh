#!/bin/bash
# kernel tests + GEMM sweep + parity + bench, one GPU call
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_kernels.py -q -x 2>&1 | tail -3
timeout 200 python tools/gemm_sweep.py ${SWEEP_CFGS:-1160 1224 1256} 2>&1 | tail -30
TRACE_HOT=1 timeout 200 python tools/gemm_trace.py 2>&1 | grep -A8 "lin1_gelu"
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x 2>&1 | tail -3
SAMRS_STREAMS=2 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench.err | tee gpurun_out/bench_last.json | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('streams', d['config'].get('tiles_in_flight_per_gpu'), 'value', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'roof', round(d['roofline']['achieved'],1), d['single_tile_in_flight'])"
tail -3 gpurun_out/bench.err

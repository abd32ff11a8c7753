#!/bin/bash
# A/B of engine variants through env switches: value + per-category ms
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -x -k "attention" 2>&1 | tail -2
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x 2>&1 | tail -2
run() { timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); k=d['single_tile_in_flight']['ms_per_step_by_kernel']; print('$1', 'value', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'single', round(d['single_tile_in_flight']['ms_per_step'],3), {a: round(b,3) for a,b in k.items()})" || tail -3 gpurun_out/bench.err; }
run "default    "
SAMRS_LN_ROWS=1 run "ln_rows    "
run "default(2) "

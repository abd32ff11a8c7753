#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_rle.py tests/test_gpu_parity.py -q -m gpu -k "rle or dropin" 2>&1 | tail -2
timeout 100 python tools/rle_bench.py 2>&1 | tail -4
run() { timeout 300 python bench.py --steps ${STEPS:-10} --warmup 3 --no-cpu-baseline 2>gpurun_out/bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); k=d['single_tile_in_flight']['ms_per_step_by_kernel']; print('$1', 'value', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'e2e ms', round(d['e2e']['ms_per_step'],3), 'single', round(d['single_tile_in_flight']['ms_per_step'],3))" || tail -3 gpurun_out/bench.err; }
run "steps10"
STEPS=30 run "steps30"

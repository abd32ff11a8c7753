#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -x -k "attention" 2>&1 | tail -3
python tools/attn_trace.py 2>&1 | grep -E "=="
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x 2>&1 | tail -3
SAMRS_STREAMS=2 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('streams', d['config'].get('tiles_in_flight_per_gpu'), 'value', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'roof', round(d['roofline']['achieved'],1), d['single_tile_in_flight'])"
tail -3 gpurun_out/bench.err

#!/bin/bash
mkdir -p gpurun_out
for ns in 1 2 3; do
  SAMRS_STREAMS=$ns timeout 600 python bench.py --steps 12 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('streams', d['config'].get('tiles_in_flight_per_gpu'), 'value', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d['ms_per_step_by_kernel'])"
done

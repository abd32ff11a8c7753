#!/bin/bash
mkdir -p gpurun_out
run() { timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('$1', 'value', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'single', round(d['single_tile_in_flight']['ms_per_step'],3), 'gemm', d['single_tile_in_flight']['ms_per_step_by_kernel']['gemm_tc'])" || tail -3 gpurun_out/bench.err; }
SAMRS_STREAMS=2 run "s2 default"
SAMRS_STREAMS=2 SAMRS_BN=256,224,256 run "s2 256,224,256"
SAMRS_STREAMS=2 SAMRS_BN=256,256,256 run "s2 256,256,256"
SAMRS_STREAMS=2 SAMRS_BN=224,224,256 run "s2 224,224,256"
SAMRS_STREAMS=3 run "s3 default"
SAMRS_STREAMS=3 SAMRS_BN=256,256,256 run "s3 256,256,256"

#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -x -k "gemm" 2>&1 | tail -4
python tools/gemm_sweep.py 2>&1 | grep lin1
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x 2>&1 | tail -3
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -3 gpurun_out/bench.err; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['ms_per_step_by_kernel'])"

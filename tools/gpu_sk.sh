#!/bin/bash
# stream-K A/B: kernel tests (bounded), GEMM sweep, bench with the three builds on one box
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -x -k "reduce_add or stream_k" 2>&1 | tail -5
timeout 300 python tools/gemm_sweep.py 1160 3160 3256 3128 2>&1 | grep -E "proj|lin2" > gpurun_out/gemm_sweep_sk.txt; cat gpurun_out/gemm_sweep_sk.txt
for lib in libsamrs_b200_nosk.so libsamrs_b200.so libsamrs_b200_sk160.so libsamrs_b200_nosk.so libsamrs_b200.so; do
  SAMRS_LIB=$lib timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/bench_sk_$lib.json 2> gpurun_out/bench_sk_$lib.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_sk_$lib.json").read().strip().splitlines()[-1])
print("$lib", "value %.0f sustained %.0f e2e %.0f frac %.3f gemm_ms %.3f" % (d["value"], d["sustained"]["value"], d["e2e"]["value"], d["roofline"]["frac"], d["single_tile_in_flight"]["ms_per_step_by_kernel"]["gemm_tc"]))
PY
done

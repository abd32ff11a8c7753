#!/bin/bash
# A/B of two library builds on one box: usage gpu_ab.sh <pytest -k expression> <lib A> <lib B> [sweep cfgs...]
mkdir -p gpurun_out
K="$1"; A="$2"; B="$3"; shift 3
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -x -k "$K" 2>&1 | tail -3
for lib in $A $B; do
  [ $# -gt 0 ] && SAMRS_LIB=$lib timeout 300 python tools/gemm_sweep.py "$@" 2>&1 | grep -E "lin1|qkv" | sed "s/^/$lib /"
done
for lib in $A $B $A $B; do
  SAMRS_LIB=$lib timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/bench_ab_$lib.json 2> gpurun_out/bench_ab_$lib.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_ab_$lib.json").read().strip().splitlines()[-1])
print("$lib", "value %.0f sustained %.0f e2e %.0f frac %.3f gemm_ms %.3f" % (d["value"], d["sustained"]["value"], d["e2e"]["value"], d["roofline"]["frac"], d["single_tile_in_flight"]["ms_per_step_by_kernel"]["gemm_tc"]))
PY
done

#!/bin/bash
# A/B of two library builds on one box: decoder timing, full GPU tests on the new build, bench with both
mkdir -p gpurun_out
A="$1"; B="$2"
for lib in $A $B; do SAMRS_LIB=$lib timeout 300 python tools/decode_ab.py 2>&1 | grep decode; done
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
for lib in $A $B $A $B; do
  SAMRS_LIB=$lib timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/bench_ab_$lib.json 2> gpurun_out/bench_ab_$lib.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_ab_$lib.json").read().strip().splitlines()[-1])
print("$lib", "value %.0f sustained %.0f e2e %.0f frac %.3f graph_replay_ms %.3f" % (d["value"], d["sustained"]["value"], d["e2e"]["value"], d["roofline"]["frac"], d["single_tile_in_flight"]["graph_replay_ms_per_step"]))
PY
done

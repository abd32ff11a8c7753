#!/bin/bash
# A/B of several library builds on ONE box (boxes differ by ~3 %): decoder timing in isolation, then the bench twice per build, interleaved.
# usage: gpurun -- 'bash tools/gpu_ab.sh libsamrs_b200_prev.so libsamrs_b200.so ...'   (builds selected through SAMRS_LIB)
mkdir -p gpurun_out
for lib in "$@"; do SAMRS_LIB=$lib timeout 300 python tools/decode_ab.py 2>&1 | grep decode | tr '\n' ' '; echo; done
for round in 1 2; do
for lib in "$@"; do
  SAMRS_LIB=$lib timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/bench_ab_$lib.json 2> gpurun_out/bench_ab_$lib.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_ab_$lib.json").read().strip().splitlines()[-1])
print("$lib", "value %.0f sustained %.0f e2e %.0f frac %.3f graph_replay_ms %.3f" % (d["value"], d["sustained"]["value"], d["e2e"]["value"], d["roofline"]["frac"], d["single_tile_in_flight"]["graph_replay_ms_per_step"]))
PY
done
done

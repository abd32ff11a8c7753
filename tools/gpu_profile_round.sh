#!/bin/bash
# round evidence: tests, smoke, bench line, ncu launch list of one step, ncu --set full of the top kernels (1 GPU)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -4 > gpurun_out/tests.log; cat gpurun_out/tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > gpurun_out/clocks_idle.csv
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/launches.csv python tools/profile_step.py > gpurun_out/ncu1.log 2>&1
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_tc2_kernel -s 9 -c 4 \
  -o gpurun_out/prof_gemm -f python tools/profile_step.py > gpurun_out/ncu2.log 2>&1
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:attn_tc2 -s 6 -c 2 \
  -o gpurun_out/prof_attn -f python tools/profile_step.py > gpurun_out/ncu3.log 2>&1
timeout 600 ncu --profile-from-start off --set full --clock-control none -k regex:"upsample4|ln_rows_stream|rle_" -c 6 \
  -o gpurun_out/prof_epi -f python tools/profile_step.py > gpurun_out/ncu4.log 2>&1
ls -la gpurun_out | head -30

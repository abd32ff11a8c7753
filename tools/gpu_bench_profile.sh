#!/bin/bash
# bench line + ncu launch list + ncu full captures of the top kernels (1 GPU)
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/launches.csv python tools/profile_step.py > gpurun_out/ncu1.log 2>&1
if [ "$1" == "full" ]; then
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_tc -s 5 -c 4 \
  -o gpurun_out/prof_gemm -f python tools/profile_step.py > gpurun_out/ncu2.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:attn_tc -s 6 -c 2 \
  -o gpurun_out/prof_attn -f python tools/profile_step.py > gpurun_out/ncu3.log 2>&1
fi
ls -la gpurun_out

#!/bin/bash
# round evidence (1 GPU): ncu launch list of one step (cold-cache, serialised), ncu --set full of the top kernels.
# gpurun brings back at most 64 MiB: the reports are reduced to their raw-metric CSV on the box; only the GEMM report is kept.
mkdir -p gpurun_out
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/launches.csv python tools/profile_step.py > gpurun_out/ncu1.log 2>&1
python tools/summarize_launches.py gpurun_out/launches.csv gpurun_out/launches_summary.csv | head -8
cap() {  # name, kernel regex, skip, count
  timeout 900 ncu --profile-from-start off --set full --clock-control none -k regex:"$2" -s $3 -c $4 \
    -o gpurun_out/prof_$1 -f python tools/profile_step.py > gpurun_out/ncu_$1.log 2>&1
  ncu -i gpurun_out/prof_$1.ncu-rep --page raw --csv > gpurun_out/prof_$1_raw.csv 2> /dev/null
  [ "$1" = "gemm" ] || rm -f gpurun_out/prof_$1.ncu-rep
}
cap gemm "gemm_tc2_kernel" 9 4
cap attn "attn_tc2" 6 2
cap epi "upsample4|ln_rows_stream|rle_pack|rle_scan|coco_string" 0 8
cap dec "t2i_attn|i2t_attn|sgemm_small|gemm_tc_kernel<128, false, 3|ln256_split|ln64_gelu" 0 8
ls -la gpurun_out | head -30; du -sh gpurun_out
# the same launch list without flushing caches between kernels (per-kernel times as they are inside a step)
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv \
  --log-file gpurun_out/launches_warm.csv python tools/profile_step.py > gpurun_out/ncu_warm.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_warm.csv gpurun_out/launches_warm_summary.csv | head -5
du -sh gpurun_out

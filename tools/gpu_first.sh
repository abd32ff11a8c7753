#!/bin/bash
# first GPU bring-up: each group in its own process so one trap does not poison the rest
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
for k in "test_gemm_fp32_out" "test_gemm_fp16_out" "test_decoder_sgemm" "test_encoder_attention and eng64 and False" "test_encoder_attention and eng64 and True" "test_encoder_attention and eng80 and False" "test_encoder_attention and eng80 and True"; do
  echo "=== $k" >> gpurun_out/kernels.log
  timeout 300 python -m pytest tests/test_gpu_kernels.py -q -x -k "$k" 2>&1 | tail -25 >> gpurun_out/kernels.log
done
echo "=== parity" >> gpurun_out/kernels.log
timeout 900 python -m pytest tests/test_gpu_parity.py -q -s 2>&1 | tail -60 >> gpurun_out/kernels.log
tail -150 gpurun_out/kernels.log

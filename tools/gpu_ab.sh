#!/bin/bash
mkdir -p gpurun_out
for lib in libsamrs_b200.so libsamrs_alt3.so libsamrs_alt4.so; do
  echo "--- $lib"
  SAMRS_LIB=$lib timeout 100 python tools/attn_trace.py 2>&1 | grep "=="
done
SAMRS_LIB=libsamrs_alt4.so timeout 200 python -m pytest tests/test_gpu_kernels.py -q -x -k attention 2>&1 | tail -1
SAMRS_LIB=libsamrs_alt3.so timeout 200 python -m pytest tests/test_gpu_kernels.py -q -x -k attention 2>&1 | tail -1

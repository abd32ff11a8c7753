#!/bin/bash
# Builds libsamrs_b200.so in-tree for sm_100a (cross-compiles without a GPU).
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
$NVCC -shared -Xcompiler -fPIC -std=c++17 -O3 -lineinfo \
  -gencode arch=compute_100a,code=sm_100a \
  -Xptxas -v \
  engine.cu -o ../libsamrs_b200.so -lcudart_static -lrt -lpthread -ldl 2> build.log || { cat build.log; exit 1; }
grep -E "error|warning" build.log | grep -v "ptxas info" | head -20 || true
echo "built $(ls -la ../libsamrs_b200.so | awk '{print $5}') bytes"

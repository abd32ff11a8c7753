#!/bin/bash
# Builds libsamrs_b200.so in-tree for sm_100a (cross-compiles without a GPU).
#   build.sh        product library
#   build.sh exp    libsamrs_b200_exp.so with -DSAMRS_EXPERIMENTS: pipeline traces, "remove one stage" switches, the
#                   TMA-multicast GEMM variant and the SAMRS_BN / SAMRS_GEMM_MCAST environment hooks (tools/ only;
#                   select it with SAMRS_LIB=libsamrs_b200_exp.so)
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
OUT=../libsamrs_b200.so
DEFS=""
if [ "$1" = "exp" ]; then OUT=../libsamrs_b200_exp.so; DEFS="-DSAMRS_EXPERIMENTS"; fi
$NVCC -shared -Xcompiler -fPIC -std=c++17 -O3 -lineinfo $DEFS $SAMRS_NVCC_FLAGS \
  -gencode arch=compute_100a,code=sm_100a \
  -Xptxas -v \
  engine.cu -o $OUT -lcudart_static -lrt -lpthread -ldl 2> build.log || { cat build.log; exit 1; }
grep -E "error|warning" build.log | grep -v "ptxas info" | head -20 || true
echo "built $(ls -la $OUT | awk '{print $5}') bytes"

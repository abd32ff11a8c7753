#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
bash tools/gpu_ab3.sh "$@"

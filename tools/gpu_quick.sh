#!/bin/bash
# Quick GPU check of the convolution kernels: tools/gpu_quick.sh <tag> [pytest -k expression]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
tag=$1; shift
expr="${1:-tensor_core_path}"
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -p no:cacheprovider -k "$expr" 2>&1 | tail -80 > gpurun_out/${tag}_quick.log
echo "quick: $(tail -1 gpurun_out/${tag}_quick.log)"; grep -E "^(FAILED|ERROR)" gpurun_out/${tag}_quick.log | head -30

#!/bin/bash
# tools/gpu_pytest.sh <tag> <pytest args...>  — runs pytest on the GPU box, keeps the tail of the log
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
tag=$1; shift
timeout 900 python -m pytest -q -p no:cacheprovider "$@" 2>&1 | tail -150 > gpurun_out/${tag}_pytest.log
echo "pytest: $(tail -1 gpurun_out/${tag}_pytest.log)"; grep -E "^(FAILED|ERROR)" gpurun_out/${tag}_pytest.log | head -30

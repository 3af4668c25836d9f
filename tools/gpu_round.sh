#!/bin/bash
# One GPU visit.  Usage: tools/gpu_round.sh <tag> [stages...]   stages: test bench launches full src
# Everything lands in gpurun_out/ (kept under 64 MiB: big .ncu-rep files are exported to CSV and removed).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
tag=$1; shift
stages="${*:-test bench}"
has() { [[ " $stages " == *" $1 "* ]]; }
if has test; then
  timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -60 > gpurun_out/${tag}_pytest.log
  echo "pytest: $(tail -1 gpurun_out/${tag}_pytest.log)"; grep -E "^(FAILED|ERROR)" gpurun_out/${tag}_pytest.log | head
fi
if has bench; then
  timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
  echo "bench rc=$?"; head -c 300 gpurun_out/${tag}_bench.json; echo
fi
if has launches; then
  export VAEGAM_CUDA_GRAPH=0      # ncu follows eager launches; the graph replays the same kernels
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv \
     --log-file gpurun_out/${tag}_launches.csv python tools/profile_step.py 3 > gpurun_out/${tag}_ncu1.log 2>&1
  echo "ncu launches rc=$?"
fi
if has full; then
  timeout 1200 ncu --set full --clock-control none -k regex:"${NCU_K:-tc_gather|wgrad_tiled|recon_|gather_kernel|bn_bwd}" \
     --launch-skip ${NCU_SKIP:-80} -c ${NCU_COUNT:-85} -o /tmp/${tag}_full python tools/profile_step.py 2 > gpurun_out/${tag}_ncu2.log 2>&1
  echo "ncu full rc=$?"
  ncu -i /tmp/${tag}_full.ncu-rep --page raw --csv > gpurun_out/${tag}_full_raw.csv 2>/dev/null
  ls -la /tmp/${tag}_full.ncu-rep gpurun_out/${tag}_full_raw.csv
fi
if has src; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"${SRC_K:-tc_gather}" \
     --launch-skip ${SRC_SKIP:-0} -c ${SRC_COUNT:-2} -o gpurun_out/${tag}_src python tools/profile_step.py 2 > gpurun_out/${tag}_ncu3.log 2>&1
  echo "ncu src rc=$?"; ls -la gpurun_out/${tag}_src.ncu-rep
fi
du -sh gpurun_out

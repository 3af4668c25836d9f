#!/bin/bash
# One GPU visit (round 2).  Usage: tools/gpu_round2.sh <tag> [stages...]   stages: test smoke bench bench4 launches
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
tag=$1; shift
stages="${*:-test smoke bench}"
has() { [[ " $stages " == *" $1 "* ]]; }
if has test; then
  timeout 2400 python -m pytest tests -m gpu -q -p no:cacheprovider -s ${PYTEST_K:+-k "$PYTEST_K"} > gpurun_out/${tag}_pytest_full.log 2>&1
  tail -120 gpurun_out/${tag}_pytest_full.log > gpurun_out/${tag}_pytest.log
  grep -E "^\[|worst|deviation" gpurun_out/${tag}_pytest_full.log > gpurun_out/${tag}_pytest_notes.log
  echo "pytest: $(tail -1 gpurun_out/${tag}_pytest.log)"; grep -E "^(FAILED|ERROR)" gpurun_out/${tag}_pytest.log | head
fi
if has smoke; then
  timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/${tag}_smoke.log
fi
if has bench; then
  timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
  echo "bench rc=$?"; head -c 400 gpurun_out/${tag}_bench.json; echo; tail -3 gpurun_out/${tag}_bench.err
fi
if has bench4; then
  timeout 900 python bench.py --config 4 --steps 10 --warmup 3 --no-extras > gpurun_out/${tag}_bench4.json 2> gpurun_out/${tag}_bench4.err
  echo "bench4 rc=$?"; head -c 400 gpurun_out/${tag}_bench4.json; echo
fi
if has launches; then
  export VAEGAM_CUDA_GRAPH=0      # ncu follows eager launches; the graph replays the same kernels
  timeout 600 python tools/profile_step.py 3 > gpurun_out/${tag}_plain.log 2>&1 &&
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv \
     --log-file gpurun_out/${tag}_launches.csv python tools/profile_step.py 3 > gpurun_out/${tag}_ncu1.log 2>&1
  echo "ncu launches rc=$?"
fi
du -sh gpurun_out

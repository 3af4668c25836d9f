#!/bin/bash
# One GPU visit: tests, bench, ncu launch list, ncu full capture of the conv / loss kernels.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
tag=${1:-run}
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/${tag}_pytest.log
echo "pytest: $(tail -1 gpurun_out/${tag}_pytest.log)"
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench rc=$?"; head -c 600 gpurun_out/${tag}_bench.json; echo
if [ "$2" != "noncu" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv \
   --log-file gpurun_out/${tag}_launches.csv python tools/profile_step.py 3 > gpurun_out/${tag}_ncu1.log 2>&1
echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'tc_gather|wgrad_tiled|recon_|gather_kernel|bn_bwd' \
   --launch-skip 80 -c 85 -o gpurun_out/${tag}_full python tools/profile_step.py 2 > gpurun_out/${tag}_ncu2.log 2>&1
echo "ncu full rc=$?"; ls -la gpurun_out/${tag}_full.ncu-rep
fi

#!/bin/bash
# One GPU session: parity tests, per-iteration times, bench, ncu launch list, ncu full capture of the NN kernel.
# Usage (from the repo root on the GPU box): bash tools/gpu_round.sh <tag>
TAG=${1:-rXX}
OUT=gpurun_out
mkdir -p $OUT
set -x
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/${TAG}_pytest.log
tail -5 $OUT/${TAG}_pytest.log
timeout 300 python tools/iter_times.py 10000000 primary > $OUT/${TAG}_iter_times.log 2>&1; tail -12 $OUT/${TAG}_iter_times.log
timeout 600 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"; cat $OUT/${TAG}_bench.json
if [ -s $OUT/${TAG}_bench.json ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
      python bench.py --no-cpu-baseline --no-e2e --steps 10 --warmup 3 > $OUT/${TAG}_ncu_list.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:nn_kernel --launch-skip 6 -c 2 \
      -o $OUT/${TAG}_nn_full -f python bench.py --no-cpu-baseline --no-e2e --steps 3 --warmup 5 > $OUT/${TAG}_ncu_full.log 2>&1
  ncu -i $OUT/${TAG}_nn_full.ncu-rep --page raw --csv > $OUT/${TAG}_nn_full_raw.csv 2>/dev/null
fi
ls -la $OUT | tail -20

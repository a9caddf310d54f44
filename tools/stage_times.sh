#!/bin/bash
# Per-kernel times of a few converged iterations (ncu launch list of a short near-regime bench).  bash tools/stage_times.sh <tag>
TAG=${1:-st}; OUT=gpurun_out; mkdir -p $OUT
ARGS="--no-cpu-baseline --no-e2e --no-regimes --regime near --steps 6 --warmup 3"
python bench.py $ARGS > $OUT/${TAG}_plain.json 2> $OUT/${TAG}_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 330 -c 120 --csv --log-file $OUT/${TAG}_launches.csv python bench.py $ARGS > $OUT/${TAG}_ncu.log 2>&1
python tools/launch_split.py $OUT/${TAG}_launches.csv | tail -6

#!/bin/bash
# ncu --set full of one kernel inside a short bench run.  bash tools/gpu_ncu.sh <tag> <kernel-regex> [skip] [bench args...]
TAG=$1; KRE=$2; SKIP=${3:-4}; shift 3
OUT=gpurun_out; mkdir -p $OUT
ARGS="--no-cpu-baseline --no-e2e --steps 3 --warmup 3 $@"
timeout 600 python bench.py $ARGS > $OUT/${TAG}_plain.json 2> $OUT/${TAG}_plain.err && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KRE --launch-skip $SKIP -c 2 \
    -o $OUT/${TAG}_full -f python bench.py $ARGS > $OUT/${TAG}_ncu.log 2>&1
cat $OUT/${TAG}_plain.json | head -c 600; echo
ncu -i $OUT/${TAG}_full.ncu-rep --page raw --csv > $OUT/${TAG}_raw.csv 2>/dev/null
ncu -i $OUT/${TAG}_full.ncu-rep --page source --csv > $OUT/${TAG}_source.csv 2>/dev/null
ls -la $OUT | grep $TAG

#!/bin/bash
# Round-end evidence on one GPU: GPU test suite, the bench line, its ncu launch list, ncu --set full of the searching-phase
# kernel and of the keep kernel in its converged state.  bash tools/gpu_final.sh <tag>
TAG=${1:-rXX}; OUT=gpurun_out; mkdir -p $OUT
timeout 1200 python -m pytest tests -x -q -m gpu > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/${TAG}_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"
A="--no-cpu-baseline --no-e2e --no-regimes --steps 20 --warmup 5"
timeout 300 python bench.py $A > $OUT/${TAG}_plain.json 2> $OUT/${TAG}_plain.err && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 330 -c 260 --csv --log-file $OUT/${TAG}_launches.csv python bench.py $A > $OUT/${TAG}_ncu_list.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nn_group_lean_kernel --launch-skip 6 -c 1 -o $OUT/${TAG}_lean_full -f python bench.py $A > $OUT/${TAG}_ncu_lean.log 2>&1
ncu -i $OUT/${TAG}_lean_full.ncu-rep --page raw --csv > $OUT/${TAG}_lean_raw.csv 2>/dev/null
ncu -i $OUT/${TAG}_lean_full.ncu-rep --page source --csv > $OUT/${TAG}_lean_source.csv 2>/dev/null
B="--no-cpu-baseline --no-e2e --no-regimes --regime near --steps 6 --warmup 3"
timeout 300 python bench.py $B > $OUT/${TAG}_plain_near.json 2> $OUT/${TAG}_plain_near.err && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nn_keep_kernel --launch-skip 6 -c 1 -o $OUT/${TAG}_keep_full -f python bench.py $B > $OUT/${TAG}_ncu_keep.log 2>&1
ncu -i $OUT/${TAG}_keep_full.ncu-rep --page raw --csv > $OUT/${TAG}_keep_raw.csv 2>/dev/null
ls -la $OUT | grep ${TAG}_ | wc -l

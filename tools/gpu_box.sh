#!/bin/bash
# Box-search bring-up: the mode-7 parity tests, then per-iteration times of modes 7 and 6.  bash tools/gpu_box.sh <tag>
TAG=${1:-bx}; OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py -x -q -m gpu -k "box or nonfinite or cli_variant or not_terrain" > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 $OUT/${TAG}_pytest.log
ICP_MODES=${ICP_MODES:-7,6} ICP_ITERS=20 timeout 600 python tools/iter_times.py 10000000 primary > $OUT/${TAG}_iter_times.log 2>&1; echo "iter_times rc=$?"; tail -30 $OUT/${TAG}_iter_times.log
ICP_MODES=7 ICP_ITERS=20 ICP_COUNT=1 timeout 600 python tools/iter_times.py 10000000 primary near > $OUT/${TAG}_iter_counts.log 2>&1; tail -30 $OUT/${TAG}_iter_counts.log

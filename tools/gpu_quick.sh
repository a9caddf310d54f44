#!/bin/bash
# Quick GPU check: parity tests + per-iteration times (no ncu).  bash tools/gpu_quick.sh <tag> [regimes...]
TAG=${1:-q}; shift
OUT=gpurun_out; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 $OUT/${TAG}_pytest.log
timeout 600 python tools/iter_times.py 10000000 ${@:-primary stress} > $OUT/${TAG}_iter_times.log 2>&1; echo "iter_times rc=$?"; tail -40 $OUT/${TAG}_iter_times.log

#!/bin/bash
# quick 1-GPU visit: ALL -m gpu tests, the default bench line
TAG=${1:-quick3}
OUT=gpurun_out/$TAG
mkdir -p $OUT
( time timeout 1500 python -m pytest tests -m gpu -q --maxfail=8 ) > $OUT/pytest_gpu.log 2>&1
echo "pytest exit: $?" >> $OUT/pytest_gpu.log
timeout 900 python bench.py --no-cpu > $OUT/bench.json 2> $OUT/bench.err
grep -E "passed|failed" $OUT/pytest_gpu.log | tail -1; cut -c1-200 $OUT/bench.json

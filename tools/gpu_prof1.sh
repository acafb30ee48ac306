#!/bin/bash
# Short GPU-box visit: parity tests, then one ncu --set full capture of two K1 P-frame launches (and optionally the bench).
TAG=${1:-run}
OUT=gpurun_out/$TAG
mkdir -p $OUT
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $OUT/pytest_gpu.log 2>&1
echo "pytest exit: $?" >> $OUT/pytest_gpu.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k1_mb_encode' --launch-skip 17 --launch-count 2 \
    -o $OUT/k1_full -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > $OUT/ncu_full.log 2>&1
[ "$2" = bench ] && timeout 600 python bench.py --no-cpu > $OUT/bench.json 2> $OUT/bench.err
tail -3 $OUT/pytest_gpu.log; tail -2 $OUT/ncu_full.log | cut -c1-300; [ -f $OUT/bench.json ] && cut -c1-400 $OUT/bench.json

#!/bin/bash
# quick 1-GPU visit: the K1-heavy parity tests, the bench line, ncu --set full of two P-frame K1 launches
TAG=${1:-quick}
OUT=gpurun_out/$TAG
mkdir -p $OUT
( time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --maxfail=5 ) > $OUT/pytest_gpu.log 2>&1
echo "pytest exit: $?" >> $OUT/pytest_gpu.log
timeout 900 python bench.py --no-cpu > $OUT/bench.json 2> $OUT/bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k1_mb_encode' --launch-skip 17 --launch-count 2 \
    -o $OUT/k1_full -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-extras > $OUT/ncu_full.log 2>&1
tail -4 $OUT/pytest_gpu.log; cut -c1-250 $OUT/bench.json

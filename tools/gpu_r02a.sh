#!/bin/bash
# round 2 visit A: parity tests, ncu --set full of two K1 P-frame launches, bench line, real-content bench
TAG=${1:-r02a}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
( time timeout 1200 python -m pytest tests -m gpu -q --maxfail=8 ) > $OUT/pytest_gpu.log 2>&1
echo "pytest exit: $?" >> $OUT/pytest_gpu.log
timeout 600 python bench.py --no-cpu > $OUT/bench.json 2> $OUT/bench.err
timeout 300 python tools/bench_real.py > $OUT/bench_real.json 2> $OUT/bench_real.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k1_mb_encode' --launch-skip 17 --launch-count 2 \
    -o $OUT/k1_full -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > $OUT/ncu_full.log 2>&1
tail -5 $OUT/pytest_gpu.log; tail -2 $OUT/ncu_full.log | cut -c1-300; cut -c1-600 $OUT/bench.json; cat $OUT/bench_real.json

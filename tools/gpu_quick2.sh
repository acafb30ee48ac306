#!/bin/bash
# quick 1-GPU visit: parity tests, bench config 4 and config 2 (I-only), ncu of the I-frame K1 launch
TAG=${1:-quick2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
( time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --maxfail=5 ) > $OUT/pytest_gpu.log 2>&1
echo "pytest exit: $?" >> $OUT/pytest_gpu.log
timeout 900 python bench.py --no-cpu --no-extras > $OUT/bench.json 2> $OUT/bench.err
timeout 900 python bench.py --no-cpu --no-extras --config 2 > $OUT/bench_c2.json 2> $OUT/bench_c2.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k1_mb_encode' --launch-skip 16 --launch-count 2 \
    -o $OUT/k1_i -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-extras > $OUT/ncu_full.log 2>&1
grep -E "passed|failed" $OUT/pytest_gpu.log | tail -1; cut -c1-200 $OUT/bench.json; cut -c1-200 $OUT/bench_c2.json

#!/bin/bash
# round 2 visit B: parity tests, new bench line (value / value_to_host / parity / e2e / file_to_file / real_content), ncu of K1-P
TAG=${1:-r02b}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
nproc > $OUT/nproc.txt; df -h /dev/shm /tmp >> $OUT/nproc.txt 2>&1
( time timeout 1200 python -m pytest tests -m gpu -q --maxfail=8 ) > $OUT/pytest_gpu.log 2>&1
echo "pytest exit: $?" >> $OUT/pytest_gpu.log
timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k1_mb_encode' --launch-skip 17 --launch-count 2 \
    -o $OUT/k1_full -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-extras > $OUT/ncu_full.log 2>&1
tail -5 $OUT/pytest_gpu.log; tail -3 $OUT/bench.err; cat $OUT/bench.json

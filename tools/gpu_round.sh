#!/bin/bash
# One GPU-box visit: parity tests, bench line, ncu launch list of the bench command, one ncu --set full
# capture of a whole step (K1 x16, K2 x2, K3, K4).  Outputs under gpurun_out/<tag>/.
TAG=${1:-run}
MODE=${2:-full}     # quick = tests + bench only
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $OUT/pytest_gpu.log 2>&1
echo "pytest exit: $?" >> $OUT/pytest_gpu.log
timeout 600 python bench.py > $OUT/bench.json 2> $OUT/bench.err
if [ "$MODE" = full ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k[1-4]_' -c 400 --csv \
    --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > $OUT/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k[1-4]_' --launch-skip 21 --launch-count 21 \
    -o $OUT/step_full -f python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > $OUT/ncu_full.log 2>&1
fi
ls -la $OUT
tail -3 $OUT/pytest_gpu.log
cat $OUT/bench.json

#!/bin/bash
# round 2 visit F (1 GPU): parity tests, bench, ncu --set full of a whole step (K1 I+P, K2 x2, K3, zero, K4), launch list, sanitizers
TAG=${1:-r02f}
OUT=gpurun_out/$TAG
mkdir -p $OUT
( time timeout 1500 python -m pytest tests -m gpu -q --maxfail=8 ) > $OUT/pytest_gpu.log 2>&1
echo "pytest exit: $?" >> $OUT/pytest_gpu.log
timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k[1-4]_|k_zero' -c 200 --csv \
    --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-extras > $OUT/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k[1-4]_|k_zero' --launch-skip 22 --launch-count 22 \
    -o $OUT/step_full -f python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-extras > $OUT/ncu_full.log 2>&1
if [ "$2" = sanitize ]; then
K="golden or clip_classes or parameter_grid or mid_frame or word_interface or gop_sharding or regrows or async_chunks or zero_frame"
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool python -m pytest tests/test_gpu_parity.py tests/test_host_engine_gpu.py -m gpu -q -x -k "$K" > $OUT/sanitizer_$tool.txt 2>&1
done
fi
tail -4 $OUT/pytest_gpu.log; cat $OUT/bench.json | cut -c1-300; for t in memcheck racecheck synccheck; do tail -3 $OUT/sanitizer_$t.txt 2>/dev/null; done

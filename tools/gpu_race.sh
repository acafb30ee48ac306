#!/bin/bash
# racecheck + synccheck (K1-heavy subset), then ALL -m gpu tests
OUT=gpurun_out/${1:-race}
mkdir -p $OUT
K="golden or clip_classes or parameter_grid or mid_frame or word_interface or gop_sharding or regrows or async_chunks or zero_frame"
for tool in racecheck synccheck memcheck; do
timeout 600 compute-sanitizer --tool $tool python -m pytest tests/test_gpu_parity.py tests/test_host_engine_gpu.py -m gpu -q -x -k "$K" > $OUT/sanitizer_$tool.txt 2>&1
done
timeout 900 python -m pytest tests -m gpu -q --maxfail=5 > $OUT/pytest_gpu.log 2>&1
for tool in racecheck synccheck memcheck; do tail -2 $OUT/sanitizer_$tool.txt; done; tail -2 $OUT/pytest_gpu.log

#!/bin/bash
# racecheck only (K1-heavy subset), then the parity tests
OUT=gpurun_out/${1:-race}
mkdir -p $OUT
K="golden or clip_classes or parameter_grid or mid_frame or word_interface or gop_sharding or regrows or async_chunks or zero_frame"
timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py tests/test_host_engine_gpu.py -m gpu -q -x -k "$K" > $OUT/sanitizer_racecheck.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --maxfail=5 > $OUT/pytest_gpu.log 2>&1
timeout 600 python bench.py --no-cpu --no-extras --no-e2e > $OUT/bench.json 2> $OUT/bench.err
tail -3 $OUT/sanitizer_racecheck.txt; tail -2 $OUT/pytest_gpu.log; cut -c1-160 $OUT/bench.json

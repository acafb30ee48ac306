#!/bin/bash
# compute-sanitizer over the small parity tests (RTL-written fixtures, clip classes, parameter grid, stop/push4, sharding)
OUT=gpurun_out/${1:-sanitize}
mkdir -p $OUT
SEL='golden or clip_classes or parameter_grid or mid_frame or gop_lengths or sharding'
for tool in memcheck racecheck synccheck; do
  ( time timeout 1200 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -k "$SEL" ) > $OUT/$tool.log 2>&1
  echo "exit: $?" >> $OUT/$tool.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|exit:" $OUT/$tool.log | tail -4
done

#!/bin/bash
# final 2-GPU check of the final build: all -m gpu tests (multi-device ones included), bench at N=2
OUT=gpurun_out/${1:-final2}
mkdir -p $OUT
( time timeout 1500 python -m pytest tests -m gpu -q --maxfail=8 ) > $OUT/pytest_gpu.log 2>&1
echo "pytest exit: $?" >> $OUT/pytest_gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > $OUT/bench_n2.json 2> $OUT/bench_n2.err
grep -E "passed|failed" $OUT/pytest_gpu.log | tail -1; cut -c1-200 $OUT/bench_n2.json

#!/bin/bash
# the driver's own round-end sequence on one GPU: pytest -m gpu -x, smoke(), bench.py, bench.py --impl reference
OUT=gpurun_out/${1:-final1}
mkdir -p $OUT
( time timeout 1500 python -m pytest tests/ -x -q -m gpu ) > $OUT/pytest_gpu.log 2>&1
echo "pytest exit: $?" >> $OUT/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke exit: $?" >> $OUT/smoke.log
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 ) > $OUT/bench.json 2> $OUT/bench.err
( time timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 ) > $OUT/ref.json 2> $OUT/ref.err
grep -E "passed|failed" $OUT/pytest_gpu.log | tail -1; cat $OUT/smoke.log; cut -c1-220 $OUT/bench.json; cut -c1-300 $OUT/ref.json; grep real $OUT/bench.err $OUT/ref.err

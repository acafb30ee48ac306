#!/bin/bash
# round 2 visit C (2 GPUs): all -m gpu tests (multi-device ones included), bench at N=1 and N=2, concurrent-copy probe
TAG=${1:-r02c}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
nvidia-smi topo -m > $OUT/topo.txt 2>&1
nproc > $OUT/nproc.txt
( time timeout 1500 python -m pytest tests -m gpu -q --maxfail=8 ) > $OUT/pytest_gpu.log 2>&1
echo "pytest exit: $?" >> $OUT/pytest_gpu.log
timeout 900 python bench.py > $OUT/bench_n1.json 2> $OUT/bench_n1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > $OUT/bench_n2.json 2> $OUT/bench_n2.err
timeout 300 python tools/h2d_probe.py --sets 0 1 0,1 > $OUT/h2d_probe.txt 2>&1
timeout 300 python tools/h2d_probe.py --bind --sets 0,1 >> $OUT/h2d_probe.txt 2>&1
tail -6 $OUT/pytest_gpu.log; tail -3 $OUT/bench_n2.err; cat $OUT/h2d_probe.txt
python - <<'PY'
import json
for n in ('n1', 'n2'):
    try:
        d = json.load(open('gpurun_out/%s/bench_%s.json' % (''+'r02c'+'', n)))
        print(n, d['value'], d['value_to_host']['value'], d['value_to_host']['ms_per_step'], d['ms_per_step'], d['parity'], d['e2e'])
    except Exception as ex:
        print(n, 'unreadable', ex)
PY

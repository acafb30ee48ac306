#!/bin/bash
# round 2 visit E (N GPUs, N = 4 or 8): concurrent-copy probe over device sets, multi-device tests, bench at N (config 4; config 5 at N=8)
N=${1:-4}
TAG=${2:-r02e_n$N}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=index,name,pci.bus_id,clocks.sm,clocks.max.sm --format=csv > $OUT/smi.txt 2>&1
nvidia-smi topo -m > $OUT/topo.txt 2>&1
nproc > $OUT/nproc.txt; lscpu | grep -i "numa\|socket\|model name" >> $OUT/nproc.txt
if [ "$N" = 8 ]; then SETS="0 0,1,2,3 4,5,6,7 0,2,4,6 0,1,4,5 0,1,2,3,4,5,6,7"; else SETS="0 0,1 2,3 0,2 0,1,2,3"; fi
timeout 300 python tools/h2d_probe.py --sets $SETS > $OUT/copy_probe.txt 2>&1
timeout 200 python tools/h2d_probe.py --bind --sets $(echo $SETS | awk '{print $NF}') >> $OUT/copy_probe.txt 2>&1
( time timeout 900 python -m pytest tests/test_host_engine_gpu.py -m gpu -q --maxfail=8 ) > $OUT/pytest_engine.log 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N > $OUT/bench_c4.json 2> $OUT/bench_c4.err
if [ "$N" = 8 ]; then
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --config 5 > $OUT/bench_c5.json 2> $OUT/bench_c5.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 --no-extras > $OUT/bench_c4_n4_on8.json 2> $OUT/bench_c4_n4_on8.err
timeout 300 python tools/multi_trace.py 8 512 > $OUT/multi_trace.txt 2>&1
fi
cat $OUT/copy_probe.txt; tail -4 $OUT/pytest_engine.log
python - $OUT <<'PY'
import json, sys, glob
for f in sorted(glob.glob(sys.argv[1] + '/bench_*.json')):
    try:
        d = json.load(open(f))
        e = d.get('e2e') or {}
        print(f.split('/')[-1], d['n_gpus'], 'value', d['value'], 'to_host', d['value_to_host']['value'], d['value_to_host']['ms_per_step'], 'vs', d['ms_per_step'], 'nccl', d['value_to_host'].get('nccl_gather_ms'),
              'parity', (d.get('parity') or {}).get('match'), (d.get('parity') or {}).get('gops_checked'), 'e2e', e.get('value'), e.get('h2d_gbs'), e.get('ceiling_gbs'), e.get('frac'), 'multi', (e.get('e2e_multi') or {}).get('value'))
    except Exception as ex:
        print(f, 'unreadable', ex)
PY

#!/bin/bash
# round 2 visit M (8 GPUs): config 5 with the partial-GOP fix of the to-host schedule; config 4 to-host with three chunks
OUT=gpurun_out/r02m
mkdir -p $OUT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --config 5 --no-e2e --no-extras > $OUT/bench_c5.json 2> $OUT/bench_c5.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --chunks 3 --tail-gops 4 --no-e2e --no-extras > $OUT/bench_c4_3chunks.json 2> $OUT/bench_c4_3chunks.err
python - <<'PY'
import json
for f in ('bench_c5', 'bench_c4_3chunks'):
    try:
        d = json.load(open('gpurun_out/r02m/%s.json' % f))
        print(f, d['value'], d['ms_per_step'], d['value_to_host']['value'], d['value_to_host']['ms_per_step'], d['value_to_host']['chunk_frames'], d['parity'])
    except Exception as ex:
        print(f, 'unreadable', ex)
PY

#!/bin/bash
# bench.py robustness: the arena in private pinned memory (single process); an arena that cannot be created
OUT=gpurun_out/${1:-benchcheck}
mkdir -p $OUT
M2V_ARENA_DIR=private timeout 600 python bench.py --no-e2e --no-extras > $OUT/bench_private.json 2> $OUT/bench_private.err
M2V_ARENA_DIR=/nonexistent timeout 600 python bench.py --no-e2e --no-extras --no-cpu > $OUT/bench_noarena.json 2> $OUT/bench_noarena.err
for f in bench_private bench_noarena; do python - $OUT/$f.json <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(sys.argv[1].split('/')[-1], d['value'], d['value_to_host'].get('value', d['value_to_host'].get('error')), (d.get('parity') or {}).get('match'), (d.get('e2e') or {}).get('value'))
except Exception as ex:
    print(sys.argv[1], 'unreadable', ex)
PY
done
tail -3 $OUT/bench_private.err

#!/bin/bash
# Parity tests, the default bench line (config 4), then the other BASELINE.json configurations and the Q_LEVEL sweep at N=1.
TAG=${1:-sweep}
OUT=gpurun_out/$TAG
mkdir -p $OUT
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $OUT/pytest_gpu.log 2>&1
echo "pytest exit: $?" >> $OUT/pytest_gpu.log
timeout 600 python bench.py > $OUT/c4_q2.json 2> $OUT/c4_q2.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/ref.json 2> $OUT/ref.err
for c in 2 3 5; do timeout 600 python bench.py --config $c --no-cpu > $OUT/c$c.json 2> $OUT/c$c.err; done
for q in 1 3 4; do timeout 600 python bench.py --q $q --no-cpu > $OUT/c4_q$q.json 2> $OUT/c4_q$q.err; done
tail -3 $OUT/pytest_gpu.log
for f in $OUT/*.json; do echo "== $f"; python - "$f" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(d.get('impl', 'ours'), d['value'], d.get('fps'), (d.get('e2e') or {}).get('value'), (d.get('roofline') or {}).get('achieved'), (d.get('roofline') or {}).get('frac'), d.get('clocks'))
except Exception as ex:
    print('unreadable', ex)
PY
done

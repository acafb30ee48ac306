#!/bin/bash
# round 2 final 1-GPU visit: all -m gpu tests, default bench line, reference arm, other configs and the Q_LEVEL sweep,
# launch list + ncu --set full of one whole step, sanitizers
TAG=${1:-r02l}
OUT=gpurun_out/$TAG
mkdir -p $OUT
( time timeout 1500 python -m pytest tests -m gpu -q --maxfail=8 ) > $OUT/pytest_gpu.log 2>&1
echo "pytest exit: $?" >> $OUT/pytest_gpu.log
timeout 900 python bench.py > $OUT/c4_q2.json 2> $OUT/c4_q2.err
( time timeout 900 python bench.py --impl reference --steps 3 --warmup 1 ) > $OUT/ref.json 2> $OUT/ref.err
for c in 2 3 5; do timeout 600 python bench.py --config $c --no-cpu --no-extras > $OUT/c$c.json 2> $OUT/c$c.err; done
for q in 1 3 4; do timeout 600 python bench.py --q $q --no-cpu --no-extras > $OUT/c4_q$q.json 2> $OUT/c4_q$q.err; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k[1-4]_|k_zero' -c 200 --csv \
    --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-extras > $OUT/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k[1-4]_|k_zero' --launch-skip 22 --launch-count 22 \
    -o $OUT/step_full -f python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-extras > $OUT/ncu_full.log 2>&1
K="golden or clip_classes or parameter_grid or mid_frame or word_interface or gop_sharding or regrows or async_chunks or zero_frame"
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool python -m pytest tests/test_gpu_parity.py tests/test_host_engine_gpu.py -m gpu -q -x -k "$K" > $OUT/sanitizer_$tool.txt 2>&1
done
tail -4 $OUT/pytest_gpu.log; for t in memcheck racecheck synccheck; do tail -2 $OUT/sanitizer_$t.txt; done
python - $OUT <<'PY'
import json, sys, glob
for f in sorted(glob.glob(sys.argv[1] + '/c*.json')) + [sys.argv[1] + '/ref.json']:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], d.get('impl', 'ours'), d['value'], d.get('fps'), d['ms_per_step'], (d.get('value_to_host') or {}).get('value'), (d.get('parity') or {}).get('match'),
              (d.get('e2e') or {}).get('value'), (d.get('roofline') or {}).get('achieved'), (d.get('roofline') or {}).get('frac'), d.get('steps'))
    except Exception as ex:
        print(f, 'unreadable', ex)
PY

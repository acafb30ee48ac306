#!/bin/bash
# round 2 visit D (2 GPUs): tests, bench N=1, ncu K1-P and K2, multi-device trace, reader sweep of the file path, bench N=2
TAG=${1:-r02d}
OUT=gpurun_out/$TAG
mkdir -p $OUT
( time timeout 1500 python -m pytest tests -m gpu -q --maxfail=8 ) > $OUT/pytest_gpu.log 2>&1
echo "pytest exit: $?" >> $OUT/pytest_gpu.log
timeout 900 python bench.py > $OUT/bench_n1.json 2> $OUT/bench_n1.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k1_mb_encode|k2_vlc' --launch-skip 21 --launch-count 22 \
    -o $OUT/step_full -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-extras > $OUT/ncu_full.log 2>&1
M2V_TRACE=1 timeout 300 python tools/multi_trace.py 2 256 > $OUT/multi_trace.txt 2>&1
python - <<'PY' > $OUT/tb_sweep.txt 2>&1
import numpy as np, subprocess, os, sys
sys.path.insert(0, '.')
import __graft_entry__ as ge
synth = ge.load_synth()
fr = synth.s1_pan(1, 16, 1920, 1152)
with open('/dev/shm/tb_sweep.yuv', 'wb') as f:
    for _ in range(16):
        fr.tofile(f)
exe = 'fpga-mpeg2-encoder_b200/m2venc_tb'
for extra in (['-dry', '-readers', '4'], ['-dry', '-readers', '8'], ['-dry', '-readers', '12'], ['-dry', '-readers', '16'], ['-dry', '-readers', '24'],
              ['-readers', '8'], ['-readers', '12'], ['-readers', '16'], ['-readers', '24'], ['-readers', '16', '-chunk', '16'], ['-readers', '16', '-chunk', '64']):
    r = subprocess.run([exe, '-XL', '7', '-YL', '7', '-P', '15'] + extra + ['/dev/shm/tb_sweep.yuv', '1920', '1152', '/dev/shm/tb_sweep.m2v'] * 2, capture_output=True, text=True)
    print(extra, [l.split('busy=')[1] for l in r.stdout.splitlines() if 'file to file' in l], flush=True)
os.unlink('/dev/shm/tb_sweep.yuv')
PY
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > $OUT/bench_n2.json 2> $OUT/bench_n2.err
tail -4 $OUT/pytest_gpu.log; cat $OUT/tb_sweep.txt; grep -v "^\[m2v" $OUT/multi_trace.txt

#!/usr/bin/env python3
"""Throughput of the hot path on REAL picture content: the reference's 1440x704 test clip (11 frames, unpacked by
`make -C oracle` into the git-ignored oracle/_ref/data), repeated as 46 closed GOPs of I+10P resident in HBM.  Same
timing as bench.py's `value` (CUDA events of the library, whole hot path); prints one JSON line.  Informational: the
contract's workload is the synthetic S1 clip (bench.py); real pictures code ~3x more levels per pixel."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import __graft_entry__ as ge

def main():
    pkg = ge.load_package()
    W, H, P, G = 1440, 704, 10, 46
    raw = np.fromfile(os.path.join(ROOT, 'oracle', '_ref', 'data', '1440x704.yuv'), dtype=np.uint8).reshape(11, 3, H, W)
    clip = torch.from_numpy(raw).cuda()
    frames = clip.repeat(G, 1, 1, 1).contiguous()
    F = frames.shape[0]
    enc = pkg.Mpeg2Encoder(XL=7, YL=6, VECTOR_LEVEL=3, Q_LEVEL=2)
    enc.set_timing(True)
    for _ in range(3):
        ptr, n = enc.encode_gops_device(frames.data_ptr(), F, 0, W // 16, H // 16, P)
    ms = [0.0] * 5
    steps = 10
    for _ in range(steps):
        ptr, n = enc.encode_gops_device(frames.data_ptr(), F, 0, W // 16, H // 16, P)
        ms = [a + b for a, b in zip(ms, enc.kernel_ms())]
    t = ms[4] / steps
    print(json.dumps({'workload': '1440x704.yuv x %d GOPs (I+10P), VECTOR_LEVEL=3 Q_LEVEL=2, %d frames resident (%.1f GB)' % (G, F, frames.numel() / 1e9),
                      'value': round(F * W * H / t / 1e3, 1), 'unit': 'Mpixel/s', 'ms_per_step': round(t, 3),
                      'phase_ms': {'k1': round(ms[0] / steps, 3), 'k2_count': round(ms[1] / steps, 3), 'k3_k4': round(ms[2] / steps, 3), 'k2_write': round(ms[3] / steps, 3)},
                      'bytes_per_pixel_out': round(n / (F * W * H), 5)}))

if __name__ == '__main__':
    main()

#!/usr/bin/env python3
"""One process, one handle, N devices (m2v_create_multi): times the streaming path per step and, with M2V_TRACE=1, makes the
workers print their pipeline events for the last step.  usage: multi_trace.py [ndev] [frames]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge

def main():
    ndev = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    Fe = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    pkg = ge.load_package(); synth = ge.load_synth()
    W, H, P = 1920, 1152, 15
    import torch
    fr = synth.s1_pan_torch(20260929, Fe, W, H, 'cuda')
    pin = pkg.PinnedArray(Fe * 3 * W * H)
    torch.from_numpy(pin.array).view(Fe, 3, H, W).copy_(fr)
    hnp = pin.array.reshape(Fe, 3, H, W)
    sink = np.empty(64 << 20, np.uint8)
    for nd in sorted({1, ndev}):
        enc = pkg.Mpeg2Encoder(XL=7, YL=7, ndev=nd)
        ts = []
        for it in range(6):
            t0 = time.perf_counter()
            enc.begin(W // 16, H // 16, P)
            t1 = time.perf_counter()
            enc.push_frames(hnp)
            t2 = time.perf_counter()
            enc.sequence_stop()
            t3 = time.perf_counter()
            n, last = enc.drain_into(sink)
            t4 = time.perf_counter()
            ts.append((t1 - t0, t2 - t1, t3 - t2, t4 - t3, t4 - t0))
        for t in ts:
            print('ndev=%d begin %.2f push %.2f stop %.2f drain %.2f total %.2f ms -> %.1f Mpixel/s' % ((nd,) + tuple(1e3 * x for x in t) + (Fe * W * H / t[4] / 1e6,)), flush=True)
        enc.close()

if __name__ == '__main__':
    main()

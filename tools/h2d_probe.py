#!/usr/bin/env python3
"""What the box can do: N concurrent plain pinned host<->device copies, no kernels, one process per GPU.
Answers "is the end-to-end number (3 B/pixel of host->device traffic per GPU) at the ceiling of the host?" for any set
of devices, e.g. the N=4 hole of round 1 (GPUs 0-3 moved 113.7 GB/s together, all eight 180.8 GB/s).

usage: h2d_probe.py [--sets 0 0,1 0,1,2,3 0,2,4,6 4,5,6,7 0,1,2,3,4,5,6,7] [--mb 256] [--seconds 1.0] [--bind]
prints one JSON line per device set: per-device and aggregate GB/s, host->device and device->host.
"""
import argparse
import json
import multiprocessing as mp
import os
import time


def bind(index):
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        masks = pynvml.nvmlDeviceGetCpuAffinity(h, ((os.cpu_count() or 1) + 63) // 64)
        cpus = {64 * i + b for i, m in enumerate(masks) for b in range(64) if (int(m) >> b) & 1} & os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


def worker(dev, mb, seconds, do_bind, bar, q):
    import torch
    ncpu = bind(dev) if do_bind else None
    torch.cuda.set_device(dev)
    n = mb << 20
    host = torch.empty(n, dtype=torch.uint8).pin_memory()
    host.fill_(1)
    d = torch.empty(n, dtype=torch.uint8, device='cuda')
    res = {}
    for name, (dst, src) in (('h2d', (d, host)), ('d2h', (host, d))):
        for _ in range(2):
            dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        bar.wait()
        t0 = time.perf_counter(); reps = 0
        while time.perf_counter() - t0 < seconds:
            for _ in range(4):
                dst.copy_(src, non_blocking=True)
            torch.cuda.synchronize()
            reps += 4
        dt = time.perf_counter() - t0
        res[name] = reps * n / dt / 1e9
        bar.wait()
    q.put((dev, res, ncpu))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--sets', nargs='*', default=['0'])
    ap.add_argument('--mb', type=int, default=256)
    ap.add_argument('--seconds', type=float, default=1.0)
    ap.add_argument('--bind', action='store_true', help='pin each process to the CPUs NVML reports as local to its GPU')
    a = ap.parse_args()
    mp.set_start_method('spawn')
    for s in a.sets:
        devs = [int(x) for x in s.split(',')]
        bar = mp.Barrier(len(devs)); q = mp.Queue()
        ps = [mp.Process(target=worker, args=(d, a.mb, a.seconds, a.bind, bar, q)) for d in devs]
        for p in ps: p.start()
        out = [q.get() for _ in ps]
        for p in ps: p.join()
        out.sort()
        print(json.dumps({'devices': devs, 'bind': a.bind,
                          'h2d_gbs': {str(d): round(r['h2d'], 1) for d, r, _ in out}, 'h2d_total_gbs': round(sum(r['h2d'] for _, r, _ in out), 1),
                          'd2h_gbs': {str(d): round(r['d2h'], 1) for d, r, _ in out}, 'd2h_total_gbs': round(sum(r['d2h'] for _, r, _ in out), 1),
                          'cpus_per_process': out[0][2]}), flush=True)


if __name__ == '__main__':
    main()

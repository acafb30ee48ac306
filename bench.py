#!/usr/bin/env python3
"""bench.py - throughput of the MPEG-2 I/P macroblock path on N B200s (one process per GPU).

Default workload = BASELINE.json configs[3], the one the metric is quoted on: 1920x1152, GOP I+15P,
VECTOR_LEVEL=3, Q_LEVEL=2 (sweep with --q), synthetic S1 "pan" clip built directly in HBM, 512 frames
(32 closed GOPs) per GPU.  Closed GOPs shard across ranks with no data-path collective; scaling is
"weak" (every rank encodes its own 512-frame block of one long sequence; value = all pixels /
max-over-ranks device time).  --config 2|3|5 select the other BASELINE.json configurations
(5 = 2048x2048 x 1000 frames, a fixed job split by GOP across the ranks: "strong").

A step = one pass of the whole hot path over the batch: K1 mb_encode (one launch per frame index in
the GOP), K2 vlc count, K3 scans, K4 headers, K2 vlc write -> body bytes in HBM.

  value   : Mpixel/s, inputs already resident in HBM, device time from the library's CUDA events on its
            launching stream (first launch -> last kernel end), max over ranks
  e2e     : same metric through the streaming C-ABI with HOST buffers (m2v_begin / m2v_push_frames /
            m2v_stop / m2v_drain): pinned host frames -> H2D -> kernels -> D2H of the stream
  roofline: K1 (dominant kernel) algorithmic bytes / its CUDA-event time vs MEASURED_PEAKS.json
  cpu_baseline / --impl reference: the CPU oracle (a port: the reference is Verilog and no simulator
            exists in the image or on the GPU box) on the host cores over a bounded sample of the workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {   # BASELINE.json configs[1..4]  (W, H, pframes_count, VECTOR_LEVEL, default frames per GPU, scaling)
    2: dict(W=640, H=480, P=0, VL=3, frames=1024, scaling='weak', name='config2: 640x480 I-only'),
    3: dict(W=1280, H=720, P=7, VL=3, frames=512, scaling='weak', name='config3: 1280x720 I+7P'),
    4: dict(W=1920, H=1152, P=15, VL=3, frames=512, scaling='weak', name='config4: 1920x1152 I+15P'),
    5: dict(W=2048, H=2048, P=15, VL=3, frames=1000, scaling='strong', name='config5: 2048x2048 I+15P, 1000 frames total'),
}


def alg_bytes_per_pixel(p):
    """algorithmic HBM bytes per pixel of K1: reads 3.0 (4:4:4 in) + 1.5 (previous recon,
    P-frames only); writes 1.5 (recon of every frame that is followed by a P-frame); GOP average."""
    return (3.0 + p * 4.5) / (p + 1), 1.5 * p / (p + 1)


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons sampled DURING the timed region: NVML in-process (a query takes well under
    a millisecond, so a 100 ms timed region still gets dozens of samples); nvidia-smi as the fallback."""
    Q = 'clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'
    BITS = (('hw_slowdown', 0x8), ('hw_thermal_slowdown', 0x40), ('sw_thermal_slowdown', 0x20), ('sw_power_cap', 0x4))

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag, self.reasons, self.max_mhz = index, [], False, set(), None
        self.active = threading.Event()                         # set only while a timed region is running
        self.query_s = []
        self.nvml = self.handle = None
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML enumerates physical devices; honour CUDA_VISIBLE_DEVICES when it is a plain index list
            vis = os.environ.get('CUDA_VISIBLE_DEVICES', '')
            ids = [int(x) for x in vis.split(',')] if vis and all(x.strip().isdigit() for x in vis.split(',')) else None
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(ids[index] if ids and index < len(ids) else index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def run(self):
        while not self.stop_flag:
            if not self.active.wait(0.01):
                continue
            if self.nvml is not None:
                try:
                    tq = time.perf_counter()
                    self.samples.append(int(self.nvml.nvmlDeviceGetClockInfo(self.handle, self.nvml.NVML_CLOCK_SM)))
                    r = int(self.nvml.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
                    for name, bit in self.BITS:
                        if r & bit:
                            self.reasons.add(name)
                    self.query_s.append(time.perf_counter() - tq)
                except Exception:
                    pass
                time.sleep(0.001)
                continue
            try:
                o = subprocess.run(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q, '--format=csv,noheader,nounits'],
                                   capture_output=True, text=True, timeout=5).stdout.strip().split(',')
                o = [x.strip() for x in o]
                if o and o[0].isdigit():
                    self.samples.append(int(o[0]))
                if len(o) > 1 and o[1].isdigit():
                    self.max_mhz = max(self.max_mhz or 0, int(o[1]))
                for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), o[2:6]):
                    if v.lower().startswith('active'):
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self):
        sm = sorted(self.samples)
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': self.max_mhz, 'reasons': sorted(self.reasons),
                'samples': len(sm), 'source': 'nvml' if self.nvml is not None else 'nvidia-smi',
                'ms_per_query': round(1e3 * sum(self.query_s) / len(self.query_s), 2) if self.query_s else None,
                'sampled': 'during the timed regions only (device-resident steps and the e2e steps)'}


def cpu_oracle_throughput(cfg, nthreads, q, steps=1):
    """GOP-parallel run of the CPU oracle on host threads (ctypes releases the GIL): one GOP of the
    workload per thread per step.  Returns (Mpixel/s, seconds of the best step, frames per step)."""
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import oracle_binding as ob
    import __graft_entry__ as ge
    synth = ge.load_synth()
    W, H, P, VL = cfg['W'], cfg['H'], cfg['P'], cfg['VL']
    gop = max(P + 1, 4 if P == 0 else 1)                        # I-only: 4 frames per thread
    clip = synth.s1_pan(20260929, gop, W, H)
    ob.lib()
    best = None
    for _ in range(steps):
        def work(i):
            ob.encode_range(clip, i * gop * (P + 1), W // 16, H // 16, P, VL=VL, Q=q)
        th = [threading.Thread(target=work, args=(i,)) for i in range(nthreads)]
        t0 = time.perf_counter()
        for t in th: t.start()
        for t in th: t.join()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    frames = nthreads * gop
    return frames * W * H / best / 1e6, best, frames


def rtl_reference_throughput(cfg, nthreads, q, steps=1, frames_per_thread=1):
    """THE REFERENCE ITSELF on the host cores: the RTL translated by oracle/vl2c.py (oracle/_ref/*.so), one module
    instance per host thread, each fed `frames_per_thread` frames of the workload clip by the testbench replay
    (the RTL takes 64 clocks per macroblock whatever the frame type, so I and P frames cost the same).  Returns
    (Mpixel/s, seconds, frames) or None when no model for these parameters is available on this box."""
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import rtl_ref_binding as rb
    import __graft_entry__ as ge
    W, H, P, VL = cfg['W'], cfg['H'], cfg['P'], cfg['VL']
    XL, YL = 7, (7 if H > 1024 else 6)
    if not rb.available(XL, YL, VL, q):
        return None
    synth = ge.load_synth()
    clip = synth.s1_pan(20260929, frames_per_thread, W, H)
    rb.lib(XL, YL, VL, q)
    best = None
    for _ in range(steps):
        insts = [rb.RtlRef(XL, YL, VL, q) for _ in range(nthreads)]
        th = [threading.Thread(target=lambda r=r: r.sequence(clip, W // 16, H // 16, P)) for r in insts]
        t0 = time.perf_counter()
        for t in th: t.start()
        for t in th: t.join()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
        for r in insts: r.close()
    frames = nthreads * frames_per_thread
    return frames * W * H / best / 1e6, best, frames


def bind_to_gpu_numa_node(index):
    """Pin this rank's host threads (and so, by first touch, its pinned staging memory) to the CPUs NVML reports as
    local to its GPU: with 8 ranks streaming 3 B/pixel each, host buffers on the wrong socket halve the H2D rate."""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get('CUDA_VISIBLE_DEVICES', '')
        ids = [int(x) for x in vis.split(',')] if vis and all(x.strip().isdigit() for x in vis.split(',')) else None
        h = pynvml.nvmlDeviceGetHandleByIndex(ids[index] if ids and index < len(ids) else index)
        masks = pynvml.nvmlDeviceGetCpuAffinity(h, ((os.cpu_count() or 1) + 63) // 64)
        cpus = {64 * i + b for i, m in enumerate(masks) for b in range(64) if (int(m) >> b) & 1} & os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


def main():
    # stdout carries the single JSON line and nothing else: whatever a library prints to file descriptor 1 (NCCL's
    # version banner, for one) is sent to stderr, and the JSON line goes to the saved descriptor at the end
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), 'w')
    os.dup2(2, 1)

    def emit(obj):
        real_stdout.write(json.dumps(obj) + '\n'); real_stdout.flush()

    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--config', type=int, default=4, choices=sorted(CONFIGS))
    ap.add_argument('--frames', type=int, default=0, help='frames per GPU (config 5: total frames); 0 = config default')
    ap.add_argument('--q', type=int, default=2, help='Q_LEVEL 1..4')
    ap.add_argument('--e2e-frames', type=int, default=256)
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-cpu', action='store_true')
    a = ap.parse_args()
    cfg = dict(CONFIGS[a.config])
    if a.frames:
        cfg['frames'] = a.frames
    W, H, P, VL = cfg['W'], cfg['H'], cfg['P'], cfg['VL']
    rank = int(os.environ.get('RANK', '0')); world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    cores = os.cpu_count() or 1
    gop = P + 1
    config = {'workload': '%s VECTOR_LEVEL=%d Q_LEVEL=%d, %d frames%s, synthetic S1 pan' %
                          (cfg['name'], VL, a.q, cfg['frames'], '/GPU' if cfg['scaling'] == 'weak' else ' total'),
              'width': W, 'height': H, 'gop': gop, 'sharding': 'closed GOPs, contiguous blocks per rank, no data-path collective',
              'l2': 'inputs (%.1f GB per GPU) larger than the 126 MB L2' % (cfg['frames'] * 3 * W * H / 1e9 / (world if cfg['scaling'] == 'strong' else 1))}

    if a.impl == 'reference':
        # The reference's own implementation is a Verilog module and neither this image nor the GPU box has a
        # Verilog simulator; the timed CPU arm is the reference RTL translated to C++ by oracle/vl2c.py
        # (oracle/_ref, kind "reference"), one instance per host thread, 1 frame each per step (bounded sample) - or, when no
        # model is present on the box, the oracle port (kind "port"), one GOP per host thread.
        if rank != 0:
            return
        nthr = min(cores, 64)                                   # more threads than this thrash the memory system (measured on the 128-core box)
        port = cpu_oracle_throughput(cfg, nthr, a.q, steps=1)
        r = rtl_reference_throughput(cfg, min(nthr, 32), a.q, steps=max(1, min(a.steps, 2)))
        if r is not None:
            v, dt, fr = r; kind = 'reference'
            sample = '%d frames (1 per host thread, %d threads, one RTL instance per thread, 64 clocks per macroblock) of the workload clip per step' % (fr, fr)
        else:
            v, dt, fr = cpu_oracle_throughput(cfg, nthr, a.q, steps=max(1, a.steps)); kind = 'port'
            sample = '%d frames (%d per host thread) of the workload clip per step' % (fr, fr // nthr)
        emit({
            'impl': 'reference', 'metric': 'Mpixel/s', 'value': round(v, 3), 'unit': 'Mpixel/s', 'n_gpus': a.gpus, 'steps': a.steps,
            'warmup': a.warmup, 'ms_per_step': round(dt * 1e3, 3), 'higher_is_better': True, 'scaling': cfg['scaling'],
            'vs_baseline': None, 'dtype': 'u8', 'data': 'synthetic', 'config': config,
            'cpu_baseline': {'value': round(v, 3), 'unit': 'Mpixel/s', 'cores': (min(nthr, 32) if kind == 'reference' else nthr), 'kind': kind, 'sample': sample,
                             'oracle_port_mpixel_s': round(port[0], 3)},
            'e2e': {'value': round(v, 3), 'unit': 'Mpixel/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'fps': round(v * 1e6 / (W * H), 2)})
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    pkg = ge.load_package(); synth = ge.load_synth()
    from fpga_mpeg2_encoder_b200 import sharding
    ncpu_local = bind_to_gpu_numa_node(local) if world > 1 else None
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    if cfg['scaling'] == 'weak':
        F = max(gop, cfg['frames'] // gop * gop)
        n0 = rank * F                                          # this rank's block of one long sequence
        total_frames = world * F
    else:
        n0, n1 = sharding.gop_partition(cfg['frames'], P, world)[rank]
        F = n1 - n0
        total_frames = cfg['frames']
    frames = torch.empty((max(F, 1), 3, H, W), dtype=torch.uint8, device=dev)
    if F:
        synth.s1_pan_torch(20260929 + rank, F, W, H, dev, out=frames)
    torch.cuda.synchronize()
    enc = pkg.Mpeg2Encoder(XL=7, YL=7, VECTOR_LEVEL=VL, Q_LEVEL=a.q)
    enc.set_timing(True)
    mbw, mbh = W // 16, H // 16

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    body_len, ptr = 0, 0
    for _ in range(a.warmup):
        if F:
            ptr, body_len = enc.encode_gops_device(frames.data_ptr(), F, n0, mbw, mbh, P)
    barrier()
    sampler = ClockSampler(local); sampler.start(); sampler.active.set()
    l0 = enc.launch_count
    dev_ms = 0.0; kms = [0.0] * 5
    t0 = time.perf_counter()
    for _ in range(a.steps):
        if F:
            ptr, body_len = enc.encode_gops_device(frames.data_ptr(), F, n0, mbw, mbh, P)
            k = enc.kernel_ms()
            kms = [x + y for x, y in zip(kms, k)]
            dev_ms += k[4]
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    sampler.active.clear()
    launches = enc.launch_count - l0
    tm = torch.tensor([dev_ms, wall_ms], dtype=torch.float64, device=dev)       # device time, max over ranks
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    ms_per_step = float(tm[0]) / a.steps
    value = total_frames * W * H / (ms_per_step * 1e-3) / 1e6

    # gather the per-rank bodies on rank 0 over NCCL (payload only, ~0.02-0.07 B/pixel) and assemble
    class _DevView:                                            # zero-copy view of the library's device buffer
        def __init__(self, p, n):
            self.__cuda_array_interface__ = {'shape': (n,), 'typestr': '|u1', 'data': (p, False), 'version': 3}
    gather_ms = 0.0
    if world > 1:
        bt = torch.as_tensor(_DevView(ptr, body_len), device=dev).clone() if body_len else torch.empty(0, dtype=torch.uint8, device=dev)
        barrier(); tg = time.perf_counter()
        bodies = sharding.gather_bodies(bt, dist, dev)
        torch.cuda.synchronize(); gather_ms = (time.perf_counter() - tg) * 1e3
        total_stream = len(sharding.assemble_stream(pkg.sequence_header(mbw, mbh), bodies, pkg.finish_stream)) if rank == 0 else 0
    else:
        total_stream = 32 * ((34 + body_len + 4) // 32 + 1)

    # ---- end-to-end through the streaming C-ABI with host buffers (every rank streams its own block) ----
    e2e = None
    if not a.no_e2e and F:
        Fe = max(gop, min(a.e2e_frames // gop * gop, F))
        host = torch.empty((Fe, 3, H, W), dtype=torch.uint8).pin_memory()
        host.copy_(frames[:Fe])
        hnp = host.numpy()
        e2 = pkg.Mpeg2Encoder(XL=7, YL=7, VECTOR_LEVEL=VL, Q_LEVEL=a.q)
        sink = np.empty(64 << 20, np.uint8)                      # the caller's stream buffer (m2v_drain copies the words into it)
        def one():
            e2.begin(mbw, mbh, P); e2.push_frames(hnp); e2.sequence_stop()
            n, last = e2.drain_into(sink)
            while not last:                                      # a stream longer than the buffer: keep pulling (the words are consumed)
                k, last = e2.drain_into(sink)
                assert k or last
                n += k
            return n
        for _ in range(2):
            nbytes = one()
        barrier()
        sampler.active.set()
        t1 = time.perf_counter()
        for _ in range(a.steps):
            nbytes = one()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t1) / a.steps
        sampler.active.clear()
        te = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        dt = float(te[0])
        e2e = {'value': round(world * Fe * W * H / dt / 1e6, 2), 'unit': 'Mpixel/s', 'h2d_bytes_per_step': world * Fe * 3 * W * H,
               'd2h_bytes_per_step': world * nbytes, 'frames_per_gpu': Fe, 'ms_per_step': round(dt * 1e3, 3),
               'api': 'm2v_begin / m2v_push_frames(pinned host) / m2v_stop / m2v_drain, one stream per rank, max over ranks',
               'host_cpus_per_rank': ncpu_local,
               'note': 'H2D of 3 B/pixel dominates; PCIe ceiling per GPU on this box is ~54 GB/s = ~18 Gpixel/s (profiles/r01_h2d_probe.txt)'}
        e2.close()
        del host

    sampler.stop_flag = True; sampler.join()
    if rank != 0:
        dist.barrier(); dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (K1 = the mb_encode launches of a step) ----
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    peak = float(peaks.get('hbm_gbs', 6650.0))
    rd, wr = alg_bytes_per_pixel(P)
    k1_ms = kms[0] / a.steps
    k1_launches = min(gop, F)
    bytes_per_launch = F * W * H * (rd + wr) / k1_launches
    achieved = bytes_per_launch / (k1_ms / k1_launches * 1e-3) / 1e9
    roofline = {'bound': 'hbm', 'kernel': 'k1_mb_encode', 'achieved': round(achieved, 2), 'peak': peak, 'unit': 'GB/s',
                'frac': round(achieved / peak, 4), 'traffic': None,
                'peak_source': 'measured (MEASURED_PEAKS.json hbm_gbs)' if 'hbm_gbs' in peaks else 'fallback 6650 GB/s',
                'alg_bytes_per_pixel': {'read': round(rd, 4), 'write': round(wr, 4)}, 'alg_bytes_per_launch': int(bytes_per_launch),
                'avg_launch_ms': round(k1_ms / k1_launches, 4), 'launches_per_step': k1_launches,
                'note': 'K1 on P-frames is bound by the integer ALU pipe (full-search SAD + transforms), not by HBM; see DESIGN.md section 3 and profiles/',
                'phase_ms_per_step': {'k1_mb_encode': round(k1_ms, 3), 'k2_vlc_count': round(kms[1] / a.steps, 3),
                                      'k3_scans_k4_headers': round(kms[2] / a.steps, 3), 'k2_vlc_write': round(kms[3] / a.steps, 3)}}
    try:
        prof = json.load(open(os.path.join(ROOT, 'profiles', 'k1_traffic.json')))
        roofline['traffic'] = prof.get('dram_bytes_per_launch')
        # the roofline that actually binds K1: the integer ALU pipe (one warp instruction per 2 clocks per SM sub-partition).
        # Instruction counts per macroblock come from the committed ncu capture, time and clock are measured live.
        ap = prof.get('alu_pipe')
        clk = sampler.summary()['sm_mhz']
        if ap and clk and VL == 3 and k1_ms > 0:
            n_i = min(F, (F + gop - 1) // gop) * mbw * mbh
            n_p = F * mbw * mbh - n_i
            sms = torch.cuda.get_device_properties(dev).multi_processor_count
            alu = n_p * ap['alu_warp_instr_per_macroblock']['P'] + n_i * ap['alu_warp_instr_per_macroblock']['I']
            bound_ms = alu * ap['cycles_per_alu_warp_instr_per_smsp'] / (sms * 4) / (clk * 1e3)
            roofline['alu_pipe'] = {'bound_ms_per_step': round(bound_ms, 3), 'achieved_ms_per_step': round(k1_ms, 3), 'frac': round(bound_ms / k1_ms, 4),
                                    'alu_warp_instr_per_macroblock': ap['alu_warp_instr_per_macroblock'], 'sm_mhz': clk,
                                    'note': 'ALU-pipe issue time of the K1 launches of a step / their measured time (VECTOR_LEVEL=3 counts)'}
    except Exception:
        pass

    cpu = None
    if not a.no_cpu and world == 1:
        nthr = min(cores, 64)
        pv, pdt, pfr = cpu_oracle_throughput(cfg, nthr, a.q, steps=1)
        r = rtl_reference_throughput(cfg, min(nthr, 32), a.q, steps=1)
        if r is not None:
            cpu = {'value': round(r[0], 3), 'unit': 'Mpixel/s', 'cores': min(nthr, 32), 'kind': 'reference',
                   'sample': '%d frames (1 per host thread, %d threads; the reference RTL via oracle/vl2c.py, one instance per thread), %.1f s' % (r[2], r[2], r[1]),
                   'oracle_port_mpixel_s': round(pv, 3)}
        else:
            cpu = {'value': round(pv, 3), 'unit': 'Mpixel/s', 'cores': nthr, 'kind': 'port',
                   'sample': '%d frames (%d per host thread, GOP-parallel) of the workload clip, %.1f s' % (pfr, pfr // nthr, pdt)}

    line = {'metric': 'Mpixel/s', 'value': round(value, 2), 'unit': 'Mpixel/s', 'n_gpus': world, 'steps': a.steps, 'warmup': a.warmup,
            'ms_per_step': round(ms_per_step, 3), 'higher_is_better': True, 'scaling': cfg['scaling'], 'vs_baseline': None,
            'dtype': 'u8', 'data': 'synthetic', 'config': config, 'fps': round(value * 1e6 / (W * H), 1),
            'wall_ms_per_step': round(float(tm[1]) / a.steps, 3), 'stream_bytes': total_stream, 'gather_ms': round(gather_ms, 3),
            'bytes_per_pixel_out': round(body_len / max(F * W * H, 1), 5),
            'roofline': roofline, 'cpu_baseline': cpu, 'e2e': e2e, 'gpu_launches': launches, 'clocks': sampler.summary(),
            'published_yardstick': {'fpga_mpixel_s': 268, 'fpga_fps_1920x1152': 121, 'source': 'reference README.md:22 (Kintex-7 FPGA; context only)'}}
    emit(line)
    if world > 1:
        dist.barrier(); dist.destroy_process_group()


if __name__ == '__main__':
    main()

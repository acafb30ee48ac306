#!/usr/bin/env python3
"""bench.py - throughput of the MPEG-2 I/P macroblock path on N B200s (one process per GPU).

Default workload = BASELINE.json configs[3], the one the metric is quoted on: 1920x1152, GOP I+15P,
VECTOR_LEVEL=3, Q_LEVEL=2 (sweep with --q), synthetic S1 "pan" clip built directly in HBM, 512 frames
(32 closed GOPs) per GPU.  Closed GOPs shard across ranks with no data-path collective; scaling is
"weak" (every rank encodes its own 512 frames of one long sequence).  --config 2|3|5 select the other
BASELINE.json configurations (5 = 2048x2048 x 1000 frames, a fixed job split by GOP across the ranks: "strong").

A step = one pass of the whole hot path over the batch: K1 mb_encode (one launch per frame index in
the GOP), K2 vlc count, K3 scans, body zeroing, K4 headers, K2 vlc write -> body bytes in HBM.

  value         : Mpixel/s, inputs already resident in HBM, device time from the library's CUDA events on its
                  launching stream (first launch -> last kernel end), max over ranks
  value_to_host : same inputs, SURVEY 8(d) protocol: first launch -> last byte of the CONCATENATED stream of all ranks in
                  rank 0's host memory.  Every rank encodes its share in `--chunks` chunks dealt block-cyclically
                  (sharding.chunk_schedule) and copies each body device->host straight to its final offset of a shared
                  pinned arena (sharding.HostArena) while the next chunk is encoded; host clock between two barriers
  parity        : after the timed regions, randomly chosen GOPs of the bodies just produced (per rank, and of rank 0's
                  assembled stream) are compared with the CPU oracle
  e2e           : same metric through the streaming C-ABI with HOST buffers (m2v_begin / m2v_push_frames /
                  m2v_stop / m2v_drain): pinned host frames -> H2D -> kernels -> D2H of the stream; with the measured
                  ceiling of N concurrent plain pinned H2D copies beside it
  e2e_multi     : the same stream through ONE process and ONE handle over all N GPUs (m2v_create_multi), rank 0
  file_to_file  : csrc/m2venc_tb (C++ testbench replay) reading the clip from a file and writing the .m2v (N=1)
  real_content  : the hot path on the reference's own 1440x704 clip where it travelled (oracle/_ref/data)
  roofline      : K1 (dominant kernel) algorithmic bytes / its CUDA-event time vs MEASURED_PEAKS.json
  cpu_baseline / --impl reference: the reference RTL itself (translated by oracle/vl2c.py) on the host cores over a
                  bounded sample of the workload; the oracle port's figure beside it.
"""
import argparse
import hashlib
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
GOP_START = bytes([0, 0, 1, 0xB8])


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
                'sampled': 'during the timed regions only (device-resident steps, to-host steps and the e2e steps)'}


def _oracle():
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import oracle_binding as ob
    ob.lib()
    return ob


def cpu_oracle_throughput(cfg, nthreads, q, steps=1):
    """GOP-parallel run of the CPU oracle on host threads (ctypes releases the GIL): one GOP of the
    workload per thread per step.  Returns (Mpixel/s, seconds of the best step, frames per step)."""
    ob = _oracle()
    import __graft_entry__ as ge
    synth = ge.load_synth()
    W, H, P, VL = cfg['W'], cfg['H'], cfg['P'], cfg['VL']
    gop = max(P + 1, 4 if P == 0 else 1)                        # I-only: 4 frames per thread
    clip = synth.s1_pan(20260929, gop, W, H)
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


class RtlSample:
    """THE REFERENCE ITSELF on the host cores: the RTL translated by oracle/vl2c.py (oracle/_ref/*.so), one module instance per
    host thread.  A step feeds every instance an I-frame and a P-frame (i_pframes_count as in the workload) of a BAND of the
    workload clip: full width, `band` rows - the RTL spends 64 clocks per macroblock whatever the picture holds, so the band
    bounds the step to a few seconds without changing the work per macroblock."""

    def __init__(self, cfg, q, nthreads, band=576, frames=2):
        sys.path.insert(0, os.path.join(ROOT, 'tests'))
        import rtl_ref_binding as rb
        import __graft_entry__ as ge
        self.rb = rb
        W, H, P, VL = cfg['W'], cfg['H'], cfg['P'], cfg['VL']
        self.W, self.Hb, self.P, self.frames, self.n = W, min(H, band) // 16 * 16, P, frames, nthreads
        self.XL, self.YL = 7, (7 if H > 1024 else 6)
        self.ok = rb.available(self.XL, self.YL, VL, q)
        self.VL, self.q = VL, q
        if self.ok:
            synth = ge.load_synth()
            self.clip = synth.s1_pan(20260929, frames, W, H)[:, :, :self.Hb, :].copy()
            rb.lib(self.XL, self.YL, VL, q)

    def step(self):
        insts = [self.rb.RtlRef(self.XL, self.YL, self.VL, self.q) for _ in range(self.n)]
        th = [threading.Thread(target=lambda r=r: r.sequence(self.clip, self.W // 16, self.Hb // 16, self.P)) for r in insts]
        t0 = time.perf_counter()
        for t in th: t.start()
        for t in th: t.join()
        dt = time.perf_counter() - t0
        for r in insts: r.close()
        return dt

    @property
    def pixels(self):
        return self.n * self.frames * self.W * self.Hb

    def describe(self):
        return ('%d frames (I+P) of the top %dx%d band of the workload clip per RTL instance, %d instances = 1 per host thread, '
                '64 clocks per macroblock' % (self.frames, self.W, self.Hb, self.n))


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


def split_gops(buf):
    """byte offsets of the GOP headers (00 00 01 B8, byte aligned: RTL:2645-2656) in a body held as a uint8 numpy array"""
    import numpy as np
    b = buf
    hit = (b[:-3] == 0) & (b[1:-2] == 0) & (b[2:-1] == 1) & (b[3:] == 0xB8)
    return np.nonzero(hit)[0]


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
    ap.add_argument('--chunks', type=int, default=2, help='chunks per rank and step of the to-host pipeline')
    ap.add_argument('--tail-gops', type=int, default=8, help='GOPs of the last (small) chunk of the to-host pipeline: only its copy to the host is exposed')
    ap.add_argument('--e2e-frames', type=int, default=256)
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--no-extras', action='store_true', help='skip e2e_multi, file_to_file, real_content and the NCCL gather timing')
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
              'width': W, 'height': H, 'gop': gop, 'sharding': 'closed GOPs, whole GOPs per rank, no data-path collective',
              'l2': 'inputs (%.1f GB per GPU) larger than the 126 MB L2' % (cfg['frames'] * 3 * W * H / 1e9 / (world if cfg['scaling'] == 'strong' else 1))}

    if a.impl == 'reference':
        # The reference's own implementation is a Verilog module and neither this image nor the GPU box has a Verilog
        # simulator; the timed CPU arm is the reference RTL translated to C++ by oracle/vl2c.py (oracle/_ref, kind
        # "reference"), one instance per host thread - or, when no model is present on the box, the oracle port (kind
        # "port"), one GOP per host thread.  It runs exactly --warmup untimed and --steps timed steps.
        if rank != 0:
            return
        nthr = min(cores, 64)                                   # more threads than this thrash the memory system (measured on the 128-core box)
        port = cpu_oracle_throughput(cfg, nthr, a.q, steps=1)
        rs = RtlSample(cfg, a.q, min(nthr, 32))
        times = []
        if rs.ok:
            kind, used, sample, px = 'reference', rs.n, rs.describe(), rs.pixels
            for _ in range(a.warmup):
                rs.step()
            for _ in range(a.steps):
                times.append(rs.step())
        else:
            kind, used = 'port', nthr
            for _ in range(a.warmup):
                cpu_oracle_throughput(cfg, nthr, a.q, steps=1)
            for _ in range(a.steps):
                v, dt, fr = cpu_oracle_throughput(cfg, nthr, a.q, steps=1)
                times.append(dt)
            px = fr * W * H
            sample = '%d frames (%d per host thread) of the workload clip per step' % (fr, fr // nthr)
        dt = sum(times) / len(times)
        v = px / dt / 1e6
        emit({
            'impl': 'reference', 'metric': 'Mpixel/s', 'value': round(v, 3), 'unit': 'Mpixel/s', 'n_gpus': a.gpus, 'steps': len(times),
            'warmup': a.warmup, 'ms_per_step': round(dt * 1e3, 3), 'higher_is_better': True, 'scaling': cfg['scaling'],
            'vs_baseline': None, 'dtype': 'u8', 'data': 'synthetic', 'config': config,
            'cpu_baseline': {'value': round(v, 3), 'unit': 'Mpixel/s', 'cores': used, 'kind': kind, 'sample': sample,
                             'pixels_per_step': px, 'oracle_port_mpixel_s': round(port[0], 3)},
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
    fsz = 3 * W * H

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    # ---- A. device-resident: value ----
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
    ms_per_step = allmax(dev_ms) / a.steps
    wall_ms_per_step = allmax(wall_ms) / a.steps
    value = total_frames * W * H / (ms_per_step * 1e-3) / 1e6

    class _DevView:                                            # zero-copy view of the library's device buffer
        def __init__(self, p, n):
            self.__cuda_array_interface__ = {'shape': (n,), 'typestr': '|u1', 'data': (p, False), 'version': 3}
    body_a = torch.as_tensor(_DevView(ptr, body_len), device=dev).cpu().numpy() if body_len else np.zeros(0, np.uint8)

    # ---- B. to host: value_to_host (SURVEY 8(d) protocol) ----
    # weak scaling: the long sequence is dealt block-cyclically (chunk c of rank r = block c*world + r), so the bodies of
    # chunk row c travel to the host while row c+1 is encoded and only the last row's copy is exposed.  strong scaling
    # (config 5): contiguous block per rank in <= 2 chunks; a rank's offset needs the totals of the lower ranks, so its
    # copies start when every rank has published its last chunk.
    blockcyclic = cfg['scaling'] == 'weak'
    nrow = max(1, a.chunks if blockcyclic else min(a.chunks, 2))
    if blockcyclic:
        sched = sharding.chunk_schedule(F, P, world, nrow, a.tail_gops)[rank]
    else:
        sched = [(f0, k, n0 + f0) for (f0, k, _) in sharding.chunk_schedule(F, P, 1, nrow)[0]] if F else []
    arena_bytes = int(total_frames * W * H * max(0.25, 2.0 * body_len / max(F * W * H, 1))) + (1 << 20)
    name = 'm2v_bench_%s' % os.environ.get('MASTER_PORT', str(os.getpid()))
    # the arena lives in /dev/shm when that has the room (a container may give it 64 MB); a single process falls back to private
    # pinned memory; a box that refuses to map or pin it gets no to-host number - and still the rest of the line
    arena, arena_err = None, None
    def open_arena():
        shm = '/dev/shm'
        try:
            st_ = os.statvfs(shm)
            room = st_.f_bavail * st_.f_frsize
        except OSError:
            room = 0
        d_ = os.environ.get('M2V_ARENA_DIR') or (shm if room > arena_bytes + (256 << 20) else None)
        if d_ in (None, 'private'):                              # no shared memory with room: one process can use private pinned memory
            if world > 1:
                raise RuntimeError('/dev/shm has %d MB free, the arena needs %d MB' % (room >> 20, arena_bytes >> 20))
            return sharding.HostArena(pkg, name, arena_bytes, rank, world, directory=None)
        return sharding.HostArena(pkg, name, arena_bytes, rank, world, directory=d_)
    if rank == 0:
        try:
            arena = open_arena()
        except Exception as ex:
            arena_err = repr(ex)[:200]
    barrier()
    if rank != 0:
        try:
            arena = open_arena()
        except Exception as ex:
            arena_err = repr(ex)[:200]
    barrier()
    have_arena = allmax(1.0 if arena_err else 0.0) == 0.0
    if not have_arena and arena is not None:
        arena.close(); arena = None
    hdr = np.frombuffer(pkg.sequence_header(mbw, mbh), np.uint8)
    epoch = [0]; bar = [0]

    def to_host_step():
        """one step: all chunks of this rank, bodies to their final offsets of the arena; returns the stream length"""
        if rank == 0:
            arena.stream[:34] = hdr
        base = 34
        if sched:
            f0, k0, a0 = sched[0]
            enc.gops_submit(frames.data_ptr() + f0 * fsz, k0, a0, mbw, mbh, P, 0)
        mine, tot = [], np.zeros(world, np.int64)
        for c in range(nrow):
            nb = 0
            if c < len(sched):
                if c + 1 < len(sched):
                    f1, k1, a1 = sched[c + 1]
                    enc.gops_submit(frames.data_ptr() + f1 * fsz, k1, a1, mbw, mbh, P, (c + 1) & 1)
                nb = enc.gops_size(c & 1)                        # known after the scans; the write pass is still running
            epoch[0] += 1
            arena.publish(c, epoch[0], nb)
            sz = arena.sizes(c, epoch[0])
            tot += sz
            if blockcyclic:
                if c < len(sched):
                    enc.gops_fetch(c & 1, arena.stream_addr + base + int(sz[:rank].sum()), nb)
                base += int(sz.sum())
            else:
                mine.append(nb)
        if not blockcyclic:
            off = 34 + int(tot[:rank].sum())
            for c in range(len(sched)):
                enc.gops_fetch(c & 1, arena.stream_addr + off, mine[c])
                off += mine[c]
            base = 34 + int(tot.sum())
        if sched:
            enc.gops_wait(0)
            if len(sched) > 1:
                enc.gops_wait(1)
        bar[0] += 1
        arena.barrier(bar[0])
        if rank == 0:                                            # end code and the final padded word (RTL:2621-2628, 2932-2937)
            t = 32 * ((base + 4) // 32 + 1)
            arena.stream[base:t] = 0
            arena.stream[base + 2] = 1; arena.stream[base + 3] = 0xB7
        return base

    stream_len, th_ms, launches_to_host, value_to_host = 0, 0.0, 0, None
    if have_arena:
        for _ in range(max(2, a.warmup)):
            stream_len = to_host_step()
        barrier()
        bar[0] += 1; arena.barrier(bar[0])
        sampler.active.set()
        l1 = enc.launch_count
        th0 = time.perf_counter()
        for _ in range(a.steps):
            stream_len = to_host_step()
        th_ms = (time.perf_counter() - th0) * 1e3 / a.steps
        sampler.active.clear()
        launches_to_host = enc.launch_count - l1
        th_ms = allmax(th_ms)
        value_to_host = total_frames * W * H / (th_ms * 1e-3) / 1e6
    total_stream = 32 * ((stream_len + 4) // 32 + 1) if have_arena else 32 * ((34 + world * body_len + 4) // 32 + 1)

    # ---- C. parity of what was just produced (GOPs picked at random, against the CPU oracle) ----
    parity = None
    if not a.no_cpu and F:
        ob = _oracle()
        rng = np.random.default_rng(20261017 + rank)
        ngop_r = (F + gop - 1) // gop
        picks = sorted(int(x) for x in rng.choice(ngop_r, size=min(2, ngop_r), replace=False))
        # (1) body of the device-resident steps: this rank's contiguous block, n0 = rank * F
        pos_a = split_gops(body_a)
        ok_a = len(pos_a) == ngop_r
        jobs = []                                                # (kind, frames ndarray, absolute index, expected bytes)
        for g in picks:
            fr_g = frames[g * gop:min((g + 1) * gop, F)].cpu().numpy()
            if ok_a:
                end = int(pos_a[g + 1]) if g + 1 < ngop_r else len(body_a)
                jobs.append(('device', fr_g, n0 + g * gop, body_a[int(pos_a[g]):end].tobytes()))
            # (2) the same frames in the to-host schedule: absolute index from the chunk they sit in
            for (f0, k, a0) in sched:
                if f0 <= g * gop < f0 + k:
                    jobs.append(('to_host', fr_g, a0 + g * gop - f0, None))
        res = [None] * len(jobs)

        def work(i):
            res[i] = ob.encode_range(jobs[i][1], jobs[i][2], mbw, mbh, P, VL=VL, Q=a.q)
        th = [threading.Thread(target=work, args=(i,)) for i in range(len(jobs))]
        for t in th: t.start()
        for t in th: t.join()
        match_dev = ok_a and all(res[i] == jobs[i][3] for i in range(len(jobs)) if jobs[i][0] == 'device')
        mine = [(jobs[i][2] // gop, hashlib.sha256(res[i]).hexdigest()) for i in range(len(jobs)) if jobs[i][0] == 'to_host']
        ndev_checked = sum(1 for j in jobs if j[0] == 'device')
        gathered = [None] * world
        devflags = [None] * world
        if world > 1:
            dist.all_gather_object(gathered, mine)
            dist.all_gather_object(devflags, (bool(match_dev), ndev_checked))
        else:
            gathered, devflags = [mine], [(bool(match_dev), ndev_checked)]
        if rank == 0:
            ngop_all = (total_frames + gop - 1) // gop
            st = arena.stream[:stream_len] if have_arena else np.zeros(0, np.uint8)
            pos = split_gops(st) if have_arena else []
            ok = len(pos) == ngop_all
            checked, match_host = 0, ok or not have_arena
            if ok:
                for lst in gathered:
                    for (g_abs, h) in lst:
                        end = int(pos[g_abs + 1]) if g_abs + 1 < ngop_all else stream_len
                        match_host = match_host and hashlib.sha256(st[int(pos[g_abs]):end].tobytes()).hexdigest() == h
                        checked += 1
            parity = {'match': bool(match_host and all(f for f, _ in devflags)), 'gops_checked': checked + sum(n for _, n in devflags),
                      'device_resident_bodies': {'match': bool(all(f for f, _ in devflags)), 'gops': sum(n for _, n in devflags)},
                      'assembled_stream_on_rank0': {'match': bool(match_host), 'gops': checked, 'gop_headers_found': int(len(pos)), 'gop_headers_expected': ngop_all},
                      'against': 'oracle/m2v_oracle.c (encode_range of the same frames at the same absolute frame index)'}

    # ---- NCCL alternative of the gather (payloads to rank 0's HBM with one grouped send/recv), warmed up ----
    gather_ms = None
    if world > 1 and not a.no_extras:
        bt = torch.as_tensor(_DevView(ptr, body_len), device=dev).clone() if body_len else torch.empty(0, dtype=torch.uint8, device=dev)
        for _ in range(2):
            sharding.gather_bodies(bt, dist, dev)
        barrier(); tg = time.perf_counter()
        for _ in range(5):
            bodies = sharding.gather_bodies(bt, dist, dev)
        torch.cuda.synchronize()
        gather_ms = allmax((time.perf_counter() - tg) * 1e3 / 5)
        del bodies, bt

    # ---- D. end-to-end through the streaming C-ABI with host buffers (every rank streams its own block) ----
    e2e = None
    if not a.no_e2e and F:
        Fe = max(gop, min(a.e2e_frames // gop * gop, F))
        pin = pkg.PinnedArray(Fe * fsz)                          # m2v_alloc_host: pinned, and seen as pinned by every device of the process
        host = torch.from_numpy(pin.array).view(Fe, 3, H, W)
        host.copy_(frames[:Fe])
        hnp = pin.array.reshape(Fe, 3, H, W)
        e2 = pkg.Mpeg2Encoder(XL=7, YL=7, VECTOR_LEVEL=VL, Q_LEVEL=a.q)
        sink = np.empty(64 << 20, np.uint8)                      # the caller's stream buffer (m2v_drain copies the words into it)
        def one(e=e2, src=hnp, pushes=1):
            e.begin(mbw, mbh, P)
            for _ in range(pushes):
                e.push_frames(src)
            e.sequence_stop()
            n, last = e.drain_into(sink)
            while not last:                                      # a stream longer than the buffer: keep pulling (the words are consumed)
                k, last = e.drain_into(sink)
                assert k or last
                n += k
            return n
        for _ in range(2):
            nbytes = one()
        e2e_sha = hashlib.sha256(sink[:nbytes].tobytes()).hexdigest()
        barrier()
        sampler.active.set()
        t1 = time.perf_counter()
        for _ in range(a.steps):
            nbytes = one()
        torch.cuda.synchronize()
        dt = allmax((time.perf_counter() - t1) / a.steps)
        sampler.active.clear()
        # what the box can do: N concurrent plain pinned host->device copies of the same bytes, no kernels
        probe = torch.empty(Fe * fsz, dtype=torch.uint8, device=dev)
        flat = host.view(-1)
        for _ in range(2):
            probe.copy_(flat, non_blocking=True)
        barrier()
        tp = time.perf_counter()
        reps = 4
        for _ in range(reps):
            probe.copy_(flat, non_blocking=True)
        torch.cuda.synchronize()
        tprobe = allmax((time.perf_counter() - tp) / reps)
        ceiling_gbs = world * Fe * fsz / tprobe / 1e9
        del probe
        e2e_gbs = world * Fe * fsz / dt / 1e9
        e2e = {'value': round(world * Fe * W * H / dt / 1e6, 2), 'unit': 'Mpixel/s', 'h2d_bytes_per_step': world * Fe * fsz,
               'd2h_bytes_per_step': world * nbytes, 'frames_per_gpu': Fe, 'ms_per_step': round(dt * 1e3, 3),
               'api': 'm2v_begin / m2v_push_frames(pinned host) / m2v_stop / m2v_drain, one stream per rank, max over ranks',
               'host_cpus_per_rank': ncpu_local, 'h2d_gbs': round(e2e_gbs, 2), 'ceiling_gbs': round(ceiling_gbs, 2), 'frac': round(e2e_gbs / ceiling_gbs, 4),
               'ceiling': '%d concurrent plain pinned cudaMemcpyAsync host->device streams of the same bytes, no kernels, measured in this run' % world,
               'note': 'host->device copy of 3 B/pixel dominates'}
        e2.close()

        # ---- E. the same stream through ONE process and ONE handle over all GPUs (m2v_create_multi), rank 0 ----
        if world > 1 and not a.no_extras:
            barrier()                                             # the other ranks then wait on the HOST (arena flag): a rank parked in an NCCL
            if rank == 0:                                         # barrier keeps a kernel spinning on its GPU, which this handle also drives
                try:
                    em = pkg.Mpeg2Encoder(XL=7, YL=7, VECTOR_LEVEL=VL, Q_LEVEL=a.q, ndev=world)
                    nb = one(em)
                    same = hashlib.sha256(sink[:nb].tobytes()).hexdigest() == e2e_sha
                    for _ in range(2):
                        one(em, pushes=world)
                    tm0 = time.perf_counter()
                    for _ in range(a.steps):
                        one(em, pushes=world)                     # the same work as the N ranks together: N pushes of the clip, one sequence
                    dtm = (time.perf_counter() - tm0) / a.steps
                    em.close()
                    e2e['e2e_multi'] = {'value': round(world * Fe * W * H / dtm / 1e6, 2), 'unit': 'Mpixel/s', 'devices': world, 'processes': 1, 'ms_per_step': round(dtm * 1e3, 3),
                                        'frames': world * Fe, 'stream_equals_single_device': bool(same),
                                        'api': 'm2v_create_multi(%d) + m2v_begin / m2v_push_frames / m2v_stop / m2v_drain from ONE host thread; the other ranks idle' % world}
                except Exception as ex:
                    e2e['e2e_multi'] = {'error': str(ex)[:200]}
            if have_arena:
                bar[0] += 1; arena.barrier(bar[0], sleep=0.002 if rank else 0.0)
            barrier()

        # ---- F. file -> file with the C++ testbench replay (N=1) ----
        if world == 1 and not a.no_extras:
            exe = os.path.join(ROOT, 'fpga-mpeg2-encoder_b200', 'm2venc_tb')
            try:
                need = Fe * fsz + (64 << 20)
                d = '/dev/shm' if os.path.isdir('/dev/shm') and os.statvfs('/dev/shm').f_bavail * os.statvfs('/dev/shm').f_frsize > need else '/tmp'
                fin, fout = os.path.join(d, 'm2v_bench_%d.yuv' % os.getpid()), os.path.join(d, 'm2v_bench_%d.m2v' % os.getpid())
                hnp.tofile(fin)
                modes = {}
                for mode, extra in (('copy', []), ('pin', ['-pin'])):
                    r = subprocess.run([exe, '-XL', '7', '-YL', '7', '-VL', str(VL), '-Q', str(a.q), '-P', str(P)] + extra +
                                       [fin, str(W), str(H), fout, fin, str(W), str(H), fout], capture_output=True, text=True, timeout=300)
                    rates = [float(l.split(' Mpixel/s')[0].split()[-1]) for l in r.stdout.splitlines() if 'Mpixel/s file to file' in l]
                    same = hashlib.sha256(open(fout, 'rb').read()).hexdigest() == e2e_sha
                    modes[mode] = {'value': rates[-1] if rates else None, 'first_pass': rates[0] if rates else None, 'stream_equals_e2e': bool(same)}
                rd = subprocess.run([exe, '-XL', '7', '-YL', '7', '-P', str(P), '-dry', fin, str(W), str(H), fout], capture_output=True, text=True, timeout=300)
                dry = [float(l.split(' Mpixel/s')[0].split()[-1]) for l in rd.stdout.splitlines() if 'Mpixel/s file to file' in l]
                best = max((m for m in modes if modes[m]['value']), key=lambda m: modes[m]['value'], default=None)
                e2e['file_to_file'] = {'value': modes[best]['value'] if best else None, 'unit': 'Mpixel/s', 'mode': best, 'modes': modes, 'frames': Fe,
                                       'frac_of_e2e': round(modes[best]['value'] / e2e['value'], 4) if best else None,
                                       'stream_equals_e2e': bool(all(m['stream_equals_e2e'] for m in modes.values())), 'dir': d,
                                       'read_ahead_alone_mpixel_s': dry[-1] if dry else None,
                                       'api': 'csrc/m2venc_tb: copy = read-ahead into a ring of pinned chunks (up to 16 pread threads) -> m2v_push_frames -> m2v_drain -> write-behind; '
                                              'pin = the mapped file pinned in place chunk by chunk, no copy; second of two passes over the same file on one handle'}
                for f in (fin, fout):
                    os.unlink(f)
            except Exception as ex:
                e2e['file_to_file'] = {'error': str(ex)[:200]}
        del host, hnp
        pin.close()

    # ---- G. real picture content (the reference's own clip, where it travelled) ----
    real = None
    clip_path = os.path.join(ROOT, 'oracle', '_ref', 'data', '1440x704.yuv')
    if world == 1 and not a.no_extras and os.path.exists(clip_path):
        try:
            Wr, Hr, Pr, G = 1440, 704, 10, 46
            raw = np.fromfile(clip_path, dtype=np.uint8).reshape(11, 3, Hr, Wr)
            fr_r = torch.from_numpy(raw).to(dev).repeat(G, 1, 1, 1).contiguous()
            torch.cuda.synchronize()
            er = pkg.Mpeg2Encoder(XL=7, YL=6, VECTOR_LEVEL=3, Q_LEVEL=2)
            er.set_timing(True)
            for _ in range(3):
                er.encode_gops_device(fr_r.data_ptr(), fr_r.shape[0], 0, Wr // 16, Hr // 16, Pr)
            msr = 0.0
            for _ in range(a.steps):
                _, nr = er.encode_gops_device(fr_r.data_ptr(), fr_r.shape[0], 0, Wr // 16, Hr // 16, Pr)
                msr += er.kernel_ms()[4]
            msr /= a.steps
            real = {'mpixel_s': round(fr_r.shape[0] * Wr * Hr / msr / 1e3, 1), 'bytes_per_pixel_out': round(nr / (fr_r.shape[0] * Wr * Hr), 5), 'ms_per_step': round(msr, 3),
                    'workload': "the reference's 1440x704.yuv (11 frames) as 46 closed GOPs of I+10P resident in HBM, VECTOR_LEVEL=3 Q_LEVEL=2"}
            er.close(); del fr_r
        except Exception as ex:
            real = {'error': str(ex)[:200]}

    sampler.stop_flag = True; sampler.join()
    if arena is not None:
        arena.close()
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
                'read_only': {'achieved': round(achieved * rd / (rd + wr), 2), 'frac': round(achieved * rd / (rd + wr) / peak, 4), 'note': "north_star's HBM-read roofline: read bytes only"},
                'note': 'K1 on P-frames is bound by the integer ALU pipe (full-search SAD + transforms), not by HBM; see DESIGN.md section 3 and profiles/',
                'phase_ms_per_step': {'k1_mb_encode': round(k1_ms, 3), 'k2_vlc_count': round(kms[1] / a.steps, 3),
                                      'k3_scans_zero_k4_headers': round(kms[2] / a.steps, 3), 'k2_vlc_write': round(kms[3] / a.steps, 3)}}
    try:
        prof = json.load(open(os.path.join(ROOT, 'profiles', 'k1_traffic.json')))
        roofline['traffic'] = prof.get('dram_bytes_per_launch')
        roofline['traffic_source'] = 'constant from profiles/k1_traffic.json (%s), not measured in this run' % prof.get('build', 'ncu --set full capture')
        # the roofline that actually binds K1: the integer ALU pipe (one warp instruction per 2 clocks per SM sub-partition).
        # Instruction counts per macroblock come from the committed ncu capture, time and clock are measured live.
        ap_ = prof.get('alu_pipe')
        clk = sampler.summary()['sm_mhz']
        if ap_ and clk and VL == 3 and k1_ms > 0:
            n_i = min(F, (F + gop - 1) // gop) * mbw * mbh
            n_p = F * mbw * mbh - n_i
            sms = torch.cuda.get_device_properties(dev).multi_processor_count
            alu = n_p * ap_['alu_warp_instr_per_macroblock']['P'] + n_i * ap_['alu_warp_instr_per_macroblock']['I']
            bound_ms = alu * ap_['cycles_per_alu_warp_instr_per_smsp'] / (sms * 4) / (clk * 1e3)
            roofline['alu_pipe'] = {'bound_ms_per_step': round(bound_ms, 3), 'achieved_ms_per_step': round(k1_ms, 3), 'frac': round(bound_ms / k1_ms, 4),
                                    'alu_warp_instr_per_macroblock': ap_['alu_warp_instr_per_macroblock'], 'sm_mhz': clk,
                                    'note': 'issue efficiency of the kernel AS WRITTEN (ALU-pipe issue time of its own instruction stream / measured time), not algorithmic efficiency'}
    except Exception:
        pass

    cpu = None
    if not a.no_cpu and world == 1:
        nthr = min(cores, 64)
        pv, pdt, pfr = cpu_oracle_throughput(cfg, nthr, a.q, steps=1)
        rs = RtlSample(cfg, a.q, min(nthr, 32))
        if rs.ok:
            dt_r = rs.step()
            cpu = {'value': round(rs.pixels / dt_r / 1e6, 3), 'unit': 'Mpixel/s', 'cores': rs.n, 'kind': 'reference',
                   'sample': rs.describe() + ', %.1f s' % dt_r, 'oracle_port_mpixel_s': round(pv, 3)}
        else:
            cpu = {'value': round(pv, 3), 'unit': 'Mpixel/s', 'cores': nthr, 'kind': 'port',
                   'sample': '%d frames (%d per host thread, GOP-parallel) of the workload clip, %.1f s' % (pfr, pfr // nthr, pdt)}

    line = {'metric': 'Mpixel/s', 'value': round(value, 2), 'unit': 'Mpixel/s', 'n_gpus': world, 'steps': a.steps, 'warmup': a.warmup,
            'ms_per_step': round(ms_per_step, 3), 'higher_is_better': True, 'scaling': cfg['scaling'], 'vs_baseline': None,
            'dtype': 'u8', 'data': 'synthetic', 'config': config, 'fps': round(value * 1e6 / (W * H), 1),
            'wall_ms_per_step': round(wall_ms_per_step, 3), 'stream_bytes': total_stream,
            'bytes_per_pixel_out': round(body_len / max(F * W * H, 1), 5),
            'value_to_host': {'error': 'no shared pinned arena on this box: %s' % arena_err} if value_to_host is None else
                             {'value': round(value_to_host, 2), 'unit': 'Mpixel/s', 'ms_per_step': round(th_ms, 3), 'chunks_per_rank': len(sched), 'chunk_frames': [k for (_, k, _) in sched],
                              'gpu_launches': launches_to_host, 'fps': round(value_to_host * 1e6 / (W * H), 1),
                              'protocol': 'SURVEY 8(d): inputs resident in HBM; first launch -> last byte of the concatenated stream of all ranks in rank 0\'s host memory; '
                                          'host clock between two barriers, max over ranks',
                              'how': ('chunks dealt block-cyclically over the ranks; ' if blockcyclic else 'contiguous block per rank; ') +
                                     'every body copied device->host straight to its final offset of a shared pinned arena (N PCIe links in parallel, sizes exchanged through the arena, no collective)',
                              'nccl_gather_ms': None if gather_ms is None else round(gather_ms, 3),
                              'nccl_gather_note': None if gather_ms is None else 'alternative: all-gather of sizes + one grouped send/recv of the exact payloads to rank 0 HBM (warmed up, mean of 5); rank 0 would still have to copy %d MB to its host' % (total_stream >> 20)},
            'parity': parity, 'roofline': roofline, 'cpu_baseline': cpu, 'e2e': e2e, 'real_content': real,
            'gpu_launches': launches, 'clocks': sampler.summary(),
            'published_yardstick': {'fpga_mpixel_s': 268, 'fpga_fps_1920x1152': 121, 'source': 'reference README.md:22 (Kintex-7 FPGA; context only)'}}
    emit(line)
    if world > 1:
        dist.barrier(); dist.destroy_process_group()


if __name__ == '__main__':
    main()

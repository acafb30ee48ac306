"""CPU tests of the host logic: the C-ABI library loads and exports every symbol include/m2venc.h
declares (no compute calls without a GPU), framing helpers, the reciprocal-divide table, GOP
partitioning, the world_size-2 gloo gather, and the synthetic generators."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_abi_exports_match_header(pkg):
    hdr = open(os.path.join(ROOT, 'include', 'm2venc.h')).read()
    declared = set(re.findall(r'\b(m2v_[a-z0-9_]+)\s*\(', hdr)) - {'m2v_encoder'}
    assert declared == set(pkg.ABI_SYMBOLS)
    L = pkg.lib()
    for s in declared:
        assert getattr(L, s) is not None
    nm = subprocess.run(['nm', '-D', '--defined-only', pkg.LIB_PATH], capture_output=True, text=True).stdout
    for s in declared:
        assert re.search(r'\bT %s\b' % s, nm), s


def test_no_cpu_fallback_without_gpu(pkg):
    """the product must fail loudly when no B200 is present (no oracle / CPU fallback)"""
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    with pytest.raises(pkg.M2VError) as ei:
        pkg.Mpeg2Encoder()
    assert ei.value.code == pkg.M2V_ENODEV
    h = C.c_void_p()
    assert pkg.lib().m2v_create(3, 6, 3, 2, C.byref(h)) == pkg.M2V_EINVAL     # parameter sets of RTL:11-14
    assert pkg.lib().m2v_create(6, 6, 4, 2, C.byref(h)) == pkg.M2V_EINVAL
    assert pkg.lib().m2v_create(6, 6, 3, 5, C.byref(h)) == pkg.M2V_EINVAL


def test_product_does_not_reference_oracle():
    pk = os.path.join(ROOT, 'fpga-mpeg2-encoder_b200')
    for dp, _, fs in os.walk(pk):
        for f in fs:
            if f.endswith(('.py', '.cu', '.cuh', '.cpp', '.h', 'Makefile')):
                txt = open(os.path.join(dp, f), errors='ignore').read()
                assert 'oracle_binding' not in txt and 'm2v_oracle' not in txt and 'libm2v_oracle' not in txt, f


def test_framing_helpers_match_oracle(pkg, ob):
    for mbw, mbh in ((4, 4), (18, 13), (120, 72), (128, 128)):
        assert pkg.sequence_header(mbw, mbh) == ob.seq_header(mbw, mbh)
    for n in (34, 35, 60, 61, 62, 63, 64, 1000):
        data = bytes(range(256)) * 4
        out = pkg.finish_stream(data[:n])
        assert len(out) == ob.lib().m2v_oracle_tail_len(n) and out[:n] == data[:n]
        assert out[n:n + 4] == b'\x00\x00\x01\xb7' and not any(out[n + 4:])


def test_reciprocal_division_is_exact():
    """K1 replaces the intra-AC divide (RTL:2072) by umulhi(n, ceil(2^32/W)); exact for n < 2^15."""
    sys.path.insert(0, os.path.join(ROOT, 'tools'))
    import gen_tables as G
    n = np.arange(1 << 15, dtype=np.uint64)
    for w in sorted(set(sum(G.INTRA_Q, []))):
        r = np.uint64((0x100000000 + w - 1) // w)
        assert ((n * r) >> np.uint64(32) == n // np.uint64(w)).all(), w


def test_ac_table_occupancy_equals_the_rtl_ranges():
    """K2 decides "table code or escape" by looking the (run, level) pair up in Table B-14 and testing the entry for
    zero; the RTL decides with range comparisons (RTL:2533-2541).  Same set of pairs."""
    sys.path.insert(0, os.path.join(ROOT, 'tools'))
    import gen_tables as G
    ac = G.build_ac()
    for run in range(64):
        for m in range(2047):
            tab = (run == 0 and m < 40) or (run == 1 and m < 18) or (run == 2 and m < 5) or (run == 3 and m < 4) or \
                  (run >= 4 and ((run <= 6 and m < 3) or (run <= 16 and m < 2) or (run <= 31 and m < 1)))
            assert tab == (run < 32 and m < 40 and ac[run][m][1] != 0), (run, m)


def test_byte_simd_identities():
    """the packed-byte forms K1 uses for the RTL's mean2 (RTL:750-757) and mean4 (RTL:760-767), replayed in numpy:
    avg4(a,b) = (a|b) - (((a^b) & 0xFEFEFEFE) >> 1) per 32-bit word; mean4 from the five horizontal pair sums of six bytes,
    packed by parity (P02, P13, P24), summed over two rows, times 64 plus 0x00400040, bytes 1 and 3 interleaved"""
    rng = np.random.default_rng(3)
    edge = np.array([0, 1, 2, 127, 128, 254, 255], dtype=np.uint8)
    def words(n):
        b = rng.integers(0, 256, (n, 4), dtype=np.uint8)
        b[: 7 * 7] = np.array([[x, y, 255 - x, y] for x in edge for y in edge], dtype=np.uint8)
        return b
    a, b = words(200000), words(200000)[::-1].copy()
    wa, wb = a.view('<u4')[:, 0].astype(np.uint64), b.view('<u4')[:, 0].astype(np.uint64)
    got = ((wa | wb) - (((wa ^ wb) & np.uint64(0xFEFEFEFE)) >> np.uint64(1))) & np.uint64(0xFFFFFFFF)
    want = ((a.astype(np.uint16) + b + 1) >> 1).astype(np.uint8)
    assert (got.astype('<u4').view(np.uint8).reshape(-1, 4) == want).all()
    # mean4: rows r0, r1 of six bytes b0..b5 (pixels x-1..x+4); left diagonals use pair sums S0..S3, right ones S1..S4
    n = 100000
    r0 = rng.integers(0, 256, (n, 6), dtype=np.uint8); r1 = rng.integers(0, 256, (n, 6), dtype=np.uint8)
    r0[:49, :] = np.array([[x, y, 255, x, y, 255 - x] for x in edge for y in edge], dtype=np.uint8); r1[:49] = 255 - r0[:49] // 2
    M = np.uint64(0xFFFFFFFF)
    def parity(r):
        u = r.astype(np.uint64)
        e0, o0 = u[:, 0] | (u[:, 2] << np.uint64(16)), u[:, 1] | (u[:, 3] << np.uint64(16))      # v0 & 0x00FF00FF, prmt(v0, 0x4341)
        e1, o1 = u[:, 2] | (u[:, 4] << np.uint64(16)), u[:, 3] | (u[:, 5] << np.uint64(16))      # the same of zp
        return e0 + o0, o0 + e1, e1 + o1
    A, B = parity(r0), parity(r1)
    q = [((x + y) * np.uint64(64) + np.uint64(0x00400040)) & M for x, y in zip(A, B)]
    def perm7351(x, y):                                           # bytes (x.b1, y.b1, x.b3, y.b3)
        f = lambda v, k: (v >> np.uint64(8 * k)) & np.uint64(0xFF)
        return np.stack([f(x, 1), f(y, 1), f(x, 3), f(y, 3)], axis=1).astype(np.uint8)
    m4 = lambda i: ((r0[:, i].astype(np.uint16) + r0[:, i + 1] + r1[:, i] + r1[:, i + 1] + 1) >> 2).astype(np.uint8)
    assert (perm7351(q[0], q[1]) == np.stack([m4(0), m4(1), m4(2), m4(3)], axis=1)).all()      # DL
    assert (perm7351(q[1], q[2]) == np.stack([m4(1), m4(2), m4(3), m4(4)], axis=1)).all()      # DR


def test_packed_argmin_equals_the_rtl_tree(ob):
    """K1 replaces find_min_in_10_values (RTL:804-840, a comparison tree) by min(key*16 + rank) with the ranks
    8,9,4,5,6,7,0,1,2,3 -> 0..9 and the table 0x3210765498: same index for every input, ties included"""
    L = ob.lib()
    rank = {8: 0, 9: 1, 4: 2, 5: 3, 6: 4, 7: 5, 0: 6, 1: 7, 2: 8, 3: 9}
    def packed(v):
        m = min(v[i] * 16 + rank[i] for i in range(10))
        return (0x3210765498 >> (4 * (m & 15))) & 15
    import itertools
    for v in itertools.product((0, 1, 2), repeat=10):            # every tie pattern over three values
        assert packed(v) == L.m2v_oracle_find_min10((C.c_int * 10)(*v)), v
    rng = np.random.default_rng(9)
    for _ in range(20000):
        v = [int(x) for x in rng.integers(0, 6, 10)] if rng.random() < 0.5 else [int(x) for x in rng.integers(0, 0x30000, 10)]
        assert packed(v) == L.m2v_oracle_find_min10((C.c_int * 10)(*v)), v


def test_biased_residual_transform_identity():
    """K1 stages residuals as halfwords biased by +256 and removes 64*8*256 from the first output of the ROW pass: every
    other row of the RTL's transform matrix (RTL:102-112) sums to zero, so nothing else changes"""
    Mx = np.array([[64] * 8, [89, 75, 50, 18, -18, -50, -75, -89], [84, 35, -35, -84, -84, -35, 35, 84], [75, -18, -89, -50, 50, 89, 18, -75],
                   [64, -64, -64, 64, 64, -64, -64, 64], [50, -89, 18, 75, -75, -18, 89, -50], [35, -84, 84, -35, -35, 84, -84, 35],
                   [18, -50, 75, -89, 89, -75, 50, -18]], dtype=np.int64)
    assert (Mx[1:].sum(axis=1) == 0).all() and Mx[0].sum() == 512
    rng = np.random.default_rng(11)
    R = rng.integers(-255, 256, (2000, 8, 8)).astype(np.int64)
    rows = (R + 256) @ Mx.T                                       # row pass on the biased halfwords
    rows[:, :, 0] -= 64 * 8 * 256
    assert (rows == R @ Mx.T).all()
    assert (Mx @ rows == Mx @ (R @ Mx.T)).all()


def test_cached_bitstring_shift_copy_identity():
    """K2: the count pass keeps each macroblock's bitstring MSB-first in 32-bit words (last word left-aligned, zero padded); the write
    pass ORs word j = funnelshift_r(L_j, L_(j-1), o) into stream word W0+j, o = bit position & 31, plus one trailing word, byte-swapped
    into wire order.  Replayed in Python on random code sequences: the ORed words equal the plain concatenation of all bits."""
    rng = np.random.default_rng(13)
    def funnelshift_r(lo, hi, s):
        return (((hi << 32) | lo) >> (s & 31)) & 0xFFFFFFFF
    for trial in range(300):
        nmb = int(rng.integers(1, 12))
        start = int(rng.integers(0, 97))                           # bit position of the first macroblock (after headers)
        stream_bits, words, pos = [0] * start, {}, start
        for _ in range(nmb):
            codes = [(int(rng.integers(0, 1 << l)), int(l)) for l in rng.integers(1, 25, int(rng.integers(1, 40)))]
            # count pass (BitLocal)
            acc = n = 0; L = []
            for c, l in codes:
                acc = (acc << l) | c; n += l
                if n >= 32:
                    L.append((acc >> (n - 32)) & 0xFFFFFFFF); n -= 32; acc &= (1 << n) - 1
            total = sum(l for _, l in codes)
            if n:
                L.append((acc << (32 - n)) & 0xFFFFFFFF)
            assert len(L) == (total + 31) // 32
            # write pass
            W0, o, prev = pos >> 5, pos & 31, 0
            for j, Lj in enumerate(L):
                v = funnelshift_r(Lj, prev, o)
                words[W0 + j] = words.get(W0 + j, 0) | v
                prev = Lj
            v = funnelshift_r(0, prev, o)
            words[W0 + len(L)] = words.get(W0 + len(L), 0) | v
            for c, l in codes:
                stream_bits += [(c >> (l - 1 - i)) & 1 for i in range(l)]
            pos += total
        want = stream_bits + [0] * (-len(stream_bits) % 32)
        for w in range(len(want) // 32):
            expect = int(''.join(map(str, want[32 * w:32 * w + 32])), 2)
            assert words.get(w, 0) == expect, (trial, w)
        assert all(v == 0 for k, v in words.items() if k >= len(want) // 32)


def test_index_decode_is_exact():
    """K1 turns a drawn macroblock index into (GOP, row, column) with umulhi(n, ceil(2^32/d)), d = macroblocks per row /
    rows per frame (4..128): exact for every n below 2^25 = M2V_K1_MAX_MBS, the per-launch bound the host enforces."""
    rng = np.random.default_rng(5)
    for d in range(4, 129):
        m = np.uint64((0x100000000 + d - 1) // d)
        k = np.arange(0, (1 << 25) // d + 1, dtype=np.uint64) * np.uint64(d)          # every multiple of d, and its neighbours
        n = np.concatenate([k, k[1:] - np.uint64(1), k + np.uint64(1), rng.integers(0, 1 << 25, 200000, dtype=np.uint64),
                            np.arange((1 << 25) - 70000, 1 << 25, dtype=np.uint64)])
        n = n[n < (1 << 25)]
        assert ((n * m) >> np.uint64(32) == n // np.uint64(d)).all(), d
    src = open(os.path.join(ROOT, 'fpga-mpeg2-encoder_b200', 'csrc', 'm2v_kernels.cuh')).read()
    assert '#define M2V_K1_MAX_MBS (1l << 25)' in src


def test_gop_partition(pkg):
    from fpga_mpeg2_encoder_b200 import sharding
    for nfr, P, world in ((512, 15, 8), (1000, 15, 8), (22, 3, 3), (5, 7, 4), (16, 0, 5)):
        parts = sharding.gop_partition(nfr, P, world)
        assert len(parts) == world and parts[0][0] == 0 and parts[-1][1] == nfr
        for (a, b), (c, d) in zip(parts, parts[1:]):
            assert b == c
        for a, b in parts:
            assert a % (P + 1) == 0 or a == nfr
    assert sharding.gop_partition(1000, 15, 8)[0] == (0, 128)


_WORKER = r'''
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, 'tests'))
import __graft_entry__ as ge
pkg = ge.load_package(); synth = ge.load_synth()
from fpga_mpeg2_encoder_b200 import sharding
import oracle_binding as ob
dist.init_process_group('gloo', init_method='tcp://127.0.0.1:{port}', rank=int(sys.argv[1]), world_size=2)
rank = dist.get_rank()
fr = synth.s1_pan(3, 10, 64, 64); P = 3
n0, n1 = sharding.gop_partition(10, P, 2)[rank]
# the encoder of a rank is injected: on the CPU box the oracle stands in for the CUDA library
body = torch.frombuffer(bytearray(ob.encode_range(fr[n0:n1], n0, 4, 4, P)), dtype=torch.uint8)
bodies = sharding.gather_bodies(body, dist)
if rank == 0:
    got = sharding.assemble_stream(pkg.sequence_header(4, 4), bodies, pkg.finish_stream)
    assert got == ob.encode(fr, 4, 4, P), 'sharded stream differs'
    print('OK', len(got))
else:
    assert bodies is None
dist.destroy_process_group()
'''


def test_two_rank_gloo_gather(tmp_path):
    """world_size 2 over gloo: GOP partition -> per-rank encode -> gather -> rank-0 concatenation is
    byte-identical to the single-process stream."""
    import socket
    s = socket.socket(); s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]; s.close()
    script = tmp_path / 'worker.py'
    script.write_text(_WORKER.format(root=ROOT, port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert 'OK' in outs[0]


def test_chunk_schedule_covers_the_sequence_in_order(pkg):
    """block-cyclic deal for the pipelined gather: every frame exactly once, absolute indices in stream order (chunk row, then
    rank), whole GOPs, local ranges contiguous per rank; with a small tail chunk the last row is the short one"""
    from fpga_mpeg2_encoder_b200 import sharding
    for (F, P, world, chunks, tail) in ((512, 15, 1, 2, 0), (512, 15, 8, 2, 8), (512, 15, 4, 3, 2), (96, 7, 2, 5, 1), (16, 15, 2, 4, 0), (40, 3, 3, 2, 9)):
        gop = P + 1
        sched = sharding.chunk_schedule(F, P, world, chunks, tail)
        assert len(sched) == world and len({len(s) for s in sched}) == 1
        nabs = 0
        for c in range(len(sched[0])):
            for r in range(world):
                f0, k, a0 = sched[r][c]
                assert k % gop == 0 and k > 0 and a0 == nabs and a0 % gop == 0
                assert f0 == sum(x[1] for x in sched[r][:c])
                nabs += k
        assert nabs == world * (F // gop) * gop
        if tail and len(sched[0]) > 1 and F // gop > tail + len(sched[0]) - 2:
            assert sched[0][-1][1] == tail * gop
    # a trailing partial GOP (the end of a sequence, one rank): kept, in the last chunk
    for (F, P, chunks) in ((120, 15, 2), (8, 15, 3), (1000, 15, 2), (17, 3, 4)):
        sched = sharding.chunk_schedule(F, P, 1, chunks)[0]
        assert sum(k for _, k, _ in sched) == F and all(k > 0 for _, k, _ in sched)
        assert all(k % (P + 1) == 0 for _, k, _ in sched[:-1]) and [f for f, _, _ in sched] == [a for _, _, a in sched]


_ARENA_WORKER = r'''
import sys, os
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, 'tests'))
import numpy as np
import __graft_entry__ as ge
pkg = ge.load_package()
from fpga_mpeg2_encoder_b200 import sharding
rank, world, name = int(sys.argv[1]), 3, sys.argv[2]
arena = sharding.HostArena(pkg, name, 1 << 16, rank, world, pin=False)
arena.barrier(1)
base = 34
for step in range(3):
    for c in range(4):
        epoch = step * 4 + c + 1
        nb = 100 * (rank + 1) + 10 * c + step
        arena.publish(c, epoch, nb)
        sz = arena.sizes(c, epoch)
        assert list(sz) == [100 * (r + 1) + 10 * c + step for r in range(world)], (rank, step, c, sz)
        off = base + int(sz[:rank].sum())
        arena.stream[off:off + nb] = rank + 1                    # every rank writes its "body" at its final offset
        base += int(sz.sum())
    arena.barrier(step + 2)
    if rank == 0:
        got = arena.stream[34:base]
        want = np.concatenate([np.full(100 * (r + 1) + 10 * c + step, r + 1, np.uint8) for c in range(4) for r in range(world)])
        assert (got == want).all()
        print('ARENA OK', step, flush=True)
    arena.barrier(100 + step)
    base = 34
arena.close()
'''


def test_host_arena_three_processes(tmp_path):
    """sharding.HostArena without a GPU (pin=False): three processes publish sizes through the shared table, write their
    bodies at the offsets they derive, and rank 0 finds the concatenation in order - three steps, so that the epochs separate them"""
    name = 'm2v_cpu_test_%d' % os.getpid()
    script = tmp_path / 'arena_worker.py'
    script.write_text(_ARENA_WORKER.format(root=ROOT))
    p0 = subprocess.Popen([sys.executable, str(script), '0', name], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    import time
    for _ in range(600):
        if os.path.exists('/dev/shm/' + name) and os.path.getsize('/dev/shm/' + name) == (1 << 16) + (1 << 16):
            break
        time.sleep(0.05)
    ps = [p0] + [subprocess.Popen([sys.executable, str(script), str(r), name], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in (1, 2)]
    outs = [p.communicate(timeout=240)[0] for p in ps]
    assert all(p.returncode == 0 for p in ps), outs
    assert outs[0].count('ARENA OK') == 3


def test_synth_is_deterministic(synth):
    for name, gen in synth.GENERATORS.items():
        a = gen(42, 3, 64, 48); b = gen(42, 3, 64, 48)
        assert a.dtype == np.uint8 and a.shape == (3, 3, 48, 64) and (a == b).all(), name
    assert (synth.s2_white(1, 2, 64, 48) != synth.s2_white(2, 2, 64, 48)).any()


def test_tables_generator_is_current():
    """the committed generated headers equal what tools/gen_tables.py emits; in the build container
    they are also cross-checked against the RTL's tables"""
    sys.path.insert(0, os.path.join(ROOT, 'tools'))
    import gen_tables as G
    assert open(os.path.join(ROOT, 'oracle', 'm2v_tables.h')).read() == G.emit('', 'M2V_ORACLE_TABLES_H', 'Oracle copy (test infrastructure).')
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'check_tables_vs_rtl.py')], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours) prints exactly one JSON line on stdout with
    the contract's keys; here on the smallest configuration so that it takes seconds"""
    import json
    import subprocess
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--config', '2', '--steps', '1', '--warmup', '1'],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['unit'] == 'Mpixel/s' and d['higher_is_better'] is True and d['value'] > 0
    assert d['cpu_baseline']['kind'] in ('reference', 'port') and d['cpu_baseline']['cores'] >= 1 and d['cpu_baseline']['sample']
    assert d['e2e'] == {'value': d['value'], 'unit': 'Mpixel/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert 'workload' in d['config'] and d['vs_baseline'] is None and d['dtype'] == 'u8'

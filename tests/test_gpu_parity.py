"""Parity tests proper (-m gpu): the CUDA path, called through the C-ABI, against the oracle and the
committed golden fixture.  Bit-exact: integer / byte work, tolerance 0."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def _unpack(info):
    inter = (info & 1).astype(np.int8)
    mvx = ((info >> 8) & 0xFF).astype(np.uint8).view(np.int8)
    mvy = ((info >> 16) & 0xFF).astype(np.uint8).view(np.int8)
    cbp = ((info >> 24) & 63).astype(np.uint8)
    return inter, mvx, mvy, cbp


def _compare(pkg, ob, frames, P, VL=3, Q=2, XL=7, YL=7, partial_px4=0):
    n, _, H, W = frames.shape
    enc = pkg.Mpeg2Encoder(XL=XL, YL=YL, VECTOR_LEVEL=VL, Q_LEVEL=Q)
    got = enc.encode_sequence(frames, P, partial_px4=partial_px4)
    want, dbg = ob.encode(frames, W // 16, H // 16, P, XL=XL, YL=YL, VL=VL, Q=Q, partial_px4=partial_px4, want_dbg=True)
    if got != want:
        # localise: per-macroblock records and levels of the (single) batch
        nmb = (W // 16) * (H // 16)
        info, coefs = enc.debug_copy(n * nmb)
        inter, mvx, mvy, cbp = _unpack(info)
        msg = []
        for name, a, b in (('inter', inter, dbg['mb_inter']), ('mvx', mvx, dbg['mb_mvx']), ('mvy', mvy, dbg['mb_mvy']),
                           ('cbp', cbp, dbg['mb_cbp'])):
            bad = np.nonzero(a != b)[0]
            if bad.size:
                i = int(bad[0])
                msg.append('%s: %d diffs, first at frame %d mb %d: got %d want %d' % (name, bad.size, i // nmb, i % nmb, a[i], b[i]))
        # K1 only writes the tiles whose cbp bit is set (the others hold no level and K2 never reads them)
        coded = ((dbg['mb_cbp'].astype(np.int32)[:, None] >> (5 - np.arange(6))[None, :]) & 1).astype(bool)
        badc = np.nonzero(((coefs != dbg['coefs']).any(axis=2) & coded).any(axis=1))[0]
        if badc.size:
            i = int(badc[0])
            msg.append('coefs: %d mbs differ, first frame %d mb %d' % (badc.size, i // nmb, i % nmb))
        first = next((i for i in range(min(len(got), len(want))) if got[i] != want[i]), None)
        pytest.fail('stream mismatch: len %d vs %d, first byte diff %s; %s' % (len(got), len(want), first, '; '.join(msg)))
    enc.close()
    return got


def test_golden_fixtures_written_by_the_rtl(pkg):
    """tests/golden/*.m2v are outputs of the reference RTL itself (oracle/vl2c.py model, make_golden.py); the
    CUDA path must reproduce them without any oracle in the loop"""
    import json
    meta = json.load(open(os.path.join(GOLD, 'rtl_fixtures.json')))
    for name, m in meta.items():
        fr = np.fromfile(os.path.join(GOLD, name + '.yuv'), dtype=np.uint8).reshape(m['frames'], 3, m['H'], m['W'])
        want = open(os.path.join(GOLD, name + '.m2v'), 'rb').read()
        enc = pkg.Mpeg2Encoder(XL=m['XL'], YL=m['YL'], VECTOR_LEVEL=m['VL'], Q_LEVEL=m['Q'])
        assert enc.encode_sequence(fr, m['P'], partial_px4=m['partial_px4']) == want, name
        enc.close()


def test_cuda_equals_rtl_model_directly(pkg, synth):
    """where a prebuilt oracle/_ref model travelled to this box: CUDA vs the RTL model, no oracle involved"""
    import rtl_ref_binding as rb
    if not rb.available(7, 6, 3, 2):
        pytest.skip('no oracle/_ref model on this box')
    r = rb.RtlRef(7, 6, 3, 2)
    enc = pkg.Mpeg2Encoder(XL=7, YL=6, VECTOR_LEVEL=3, Q_LEVEL=2)
    for seed, (W, H, n, P) in enumerate([(160, 96, 6, 3), (64, 64, 9, 23), (288, 208, 3, 1)]):
        fr = synth.s1_pan(100 + seed, n, W, H)
        assert enc.encode_sequence(fr, P) == r.sequence(fr, W // 16, H // 16, P)


def test_intra_only_config2_shape(pkg, ob, synth):
    """config 2 shape (640x480, i_pframes_count=0) at a size the oracle finishes in seconds"""
    _compare(pkg, ob, synth.s1_pan(20260927, 6, 640, 480), 0)


@pytest.mark.parametrize('gen', ['S1', 'S2', 'S3', 'S4'])
def test_clip_classes(pkg, ob, synth, gen):
    fr = synth.GENERATORS[gen](20260925, 9, 160, 112)
    _compare(pkg, ob, fr, 7)


@pytest.mark.parametrize('VL', [1, 2, 3])
@pytest.mark.parametrize('Q', [1, 2, 3, 4])
def test_parameter_grid(pkg, ob, synth, VL, Q):
    fr = np.concatenate([synth.s1_pan(VL * 10 + Q, 5, 96, 80), synth.s4_edges(VL * 10 + Q, 4, 96, 80)])
    _compare(pkg, ob, fr, 3, VL=VL, Q=Q)


@pytest.mark.parametrize('P', [0, 1, 7, 23, 255])
def test_gop_lengths(pkg, ob, synth, P):
    _compare(pkg, ob, synth.s1_pan(P + 1, 27, 64, 64), P)


def test_mid_frame_stop_and_push4(pkg, ob, synth):
    fr = synth.s1_pan(77, 4, 80, 64)
    _compare(pkg, ob, fr, 7, partial_px4=333)
    _compare(pkg, ob, fr[:1], 7, partial_px4=1)


def test_push4_whole_frames_equals_bulk(pkg, synth):
    fr = synth.s1_pan(5, 2, 64, 64)
    enc = pkg.Mpeg2Encoder(XL=6, YL=6)
    a = enc.encode_sequence(fr, 3)
    enc.begin(4, 4, 3)
    for f in fr:
        yy, uu, vv = f[0].reshape(-1), f[1].reshape(-1), f[2].reshape(-1)
        for i in range(64 * 64 // 4):
            enc.push4(yy[4 * i:4 * i + 4], uu[4 * i:4 * i + 4], vv[4 * i:4 * i + 4])
    assert enc.sequence_busy
    enc.sequence_stop()
    b, last = enc.drain()
    assert last and a == b and not enc.sequence_busy


def test_back_to_back_sequences_and_clamp(pkg, ob, synth):
    """the testbench's scenario (TB:150): several sequences on one instance; plus the size clamp"""
    enc = pkg.Mpeg2Encoder(XL=4, YL=4, VECTOR_LEVEL=2, Q_LEVEL=3)
    for seed, (W, H) in enumerate([(128, 64), (64, 96), (256, 256)]):
        fr = synth.s1_pan(seed, 5, W, H)
        assert enc.encode_sequence(fr, 2) == ob.encode(fr, W // 16, H // 16, 2, XL=4, YL=4, VL=2, Q=3)
    assert enc.begin(100, 2, 0) == (16, 4)                   # clamp (RTL:985-991)
    with pytest.raises(pkg.M2VError):
        enc.push4([0] * 4, [0] * 4, [0] * 4); enc.begin(4, 4, 0)     # begin while busy
    enc.sequence_stop(); enc.drain()


def test_word_interface(pkg, ob, synth):
    fr = synth.s1_pan(9, 3, 64, 64)
    enc = pkg.Mpeg2Encoder()
    enc.begin(4, 4, 1); enc.push_frames(fr); enc.sequence_stop()
    words = []
    while True:
        w = enc.pull()
        if w is None:
            break
        words.append(w)
    assert [l for _, l in words] == [False] * (len(words) - 1) + [True]
    assert b''.join(w for w, _ in words) == ob.encode(fr, 4, 4, 1, XL=6, YL=6)


def test_gop_sharding_on_device(pkg, ob, synth):
    """bodies of GOP ranges, encoded independently from device-resident frames, concatenate to the
    whole-sequence stream (what bench.py --gpus N does across ranks)"""
    import torch
    from fpga_mpeg2_encoder_b200 import sharding
    fr = synth.s1_pan(31, 22, 96, 64)
    P = 3
    d = torch.from_numpy(fr).cuda()
    enc = pkg.Mpeg2Encoder(XL=7, YL=7)
    fsz = fr[0].size
    bodies = []
    out = np.zeros(4 << 20, np.uint8)
    for n0, n1 in sharding.gop_partition(22, P, 3):
        ln = enc.encode_gops_host(d.data_ptr() + n0 * fsz, n1 - n0, n0, 6, 4, P, out)
        bodies.append(out[:ln].tobytes())
    got = pkg.finish_stream(pkg.sequence_header(6, 4) + b''.join(bodies))
    assert got == ob.encode(fr, 6, 4, P, XL=7, YL=7)


def test_config3_and_config4_sizes_sample_gop(pkg, ob, synth):
    """full frame sizes of configs 3/4 (1280x720 I+7P, 1920x1152 I+15P): one short GOP each against
    the oracle, plus size-independent properties on a longer run: idempotence and GOP locality."""
    import torch
    for (W, H, P, n) in ((1280, 720, 7, 3), (1920, 1152, 15, 3)):
        fr = synth.s1_pan(W, n, W, H)
        _compare(pkg, ob, fr, P)
    W, H, P = 1920, 1152, 15
    fr = synth.s1_pan_torch(4, 40, W, H, 'cuda')
    enc = pkg.Mpeg2Encoder(XL=7, YL=7)
    out = np.zeros(64 << 20, np.uint8)
    la = enc.encode_gops_host(fr.data_ptr(), 40, 0, W // 16, H // 16, P, out)
    a = out[:la].tobytes()
    lb = enc.encode_gops_host(fr.data_ptr(), 40, 0, W // 16, H // 16, P, out)
    assert a == out[:lb].tobytes()                                              # idempotent
    fsz = 3 * W * H
    l1 = enc.encode_gops_host(fr.data_ptr() + 16 * fsz, 16, 16, W // 16, H // 16, P, out)
    mid = out[:l1].tobytes()
    assert mid in a                                                             # GOP 1 is a contiguous, independent unit
    host = fr[16:19].cpu().numpy()
    want = ob.encode_range(host, 16, W // 16, H // 16, P)
    l2 = enc.encode_gops_host(fr.data_ptr() + 16 * fsz, 3, 16, W // 16, H // 16, P, out)
    assert out[:l2].tobytes() == want


def test_full_gop_1920x1152_parameter_sweep(pkg, ob, synth):
    """BASELINE config 4 is a Q_LEVEL sweep at 1920x1152: one FULL 16-frame GOP for Q_LEVEL 1, 3, 4 (VECTOR_LEVEL=3) and for
    VECTOR_LEVEL 1, 2 (Q_LEVEL=2) against the oracle.  k1_mb_encode<1,true> / <2,true> have their own TMA box heights
    (18+4*VL window rows) and search ranges (RTL:1634-1715); the quantisers differ per Q_LEVEL (RTL:2065-2077)."""
    import threading
    import torch
    W, H, P = 1920, 1152, 15
    fr = synth.s1_pan(20261018, 16, W, H)
    d = torch.from_numpy(fr).cuda()
    cases = [(3, 1), (3, 3), (3, 4), (1, 2), (2, 2)]
    want = {}
    def work(c):
        want[c] = ob.encode_range(fr, 32, W // 16, H // 16, P, VL=c[0], Q=c[1])
    th = [threading.Thread(target=work, args=(c,)) for c in cases]
    for t in th: t.start()
    out = np.zeros(64 << 20, np.uint8)
    got = {}
    for VL, Q in cases:
        enc = pkg.Mpeg2Encoder(XL=7, YL=7, VECTOR_LEVEL=VL, Q_LEVEL=Q)
        ln = enc.encode_gops_host(d.data_ptr(), 16, 32, W // 16, H // 16, P, out)      # GOP 2 of a longer sequence (time code)
        got[(VL, Q)] = out[:ln].tobytes()
        enc.close()
    for t in th: t.join()
    for c in cases:
        assert got[c] == want[c], 'VECTOR_LEVEL=%d Q_LEVEL=%d: %d vs %d bytes' % (c[0], c[1], len(got[c]), len(want[c]))


def test_config5_full_gop(pkg, ob, synth):
    """config 5 (2048x2048, the XL=YL=7 maximum): one FULL I+15P GOP against the oracle, the oracle cut in two threads is
    not possible inside a GOP, so this is the slowest single oracle call of the suite (~3 s)"""
    import torch
    W = H = 2048
    P = 15
    fr = synth.s1_pan(20261019, 16, W, H)
    d = torch.from_numpy(fr).cuda()
    enc = pkg.Mpeg2Encoder(XL=7, YL=7)
    out = np.zeros(64 << 20, np.uint8)
    ln = enc.encode_gops_host(d.data_ptr(), 16, 16, W // 16, H // 16, P, out)
    enc.close()
    assert out[:ln].tobytes() == ob.encode_range(fr, 16, W // 16, H // 16, P)


def test_config5_maximum_size(pkg, ob, synth):
    """2048x2048 = the largest frame XL=YL=7 allows (config 5): I+2P against the oracle, and the clamp at that
    limit (RTL:985-991): asking for 129x129 macroblocks encodes exactly the 128x128 stream."""
    W = H = 2048
    fr = synth.s1_pan(5, 3, W, H)
    got = _compare(pkg, ob, fr, 15)
    enc = pkg.Mpeg2Encoder(XL=7, YL=7)
    mbw, mbh = enc.begin(129, 129, 15)
    assert (mbw, mbh) == (128, 128)
    enc.push_frames(fr); enc.sequence_stop()
    data, last = enc.drain(cap=64 << 20)
    assert last and data == got
    enc.close()


def test_streaming_batches_and_chunks_are_invisible(pkg, ob, synth):
    """the stream does not depend on how the library cuts the work (closed GOPs, RTL:2645-2656): the double-buffered
    multi-batch path of m2v_push_frames, the staged path (frame-by-frame pushes, trailing partial GOP, unfinished last
    frame) and the multi-chunk path of m2v_encode_gops_*, all forced to tiny sizes with m2v_set_limits"""
    W, H, P, n = 96, 64, 3, 23                                   # 5 whole GOPs + 3 frames
    fr = synth.s1_pan(77, n, W, H)
    want = ob.encode(fr, W // 16, H // 16, P, XL=6, YL=6)
    for batch in (4, 8, 12):                                     # 1, 2, 3 GOPs per batch
        enc = pkg.Mpeg2Encoder(XL=6, YL=6)
        enc.set_limits(batch_frames=batch, chunk_frames=4)
        assert enc.encode_sequence(fr, P) == want, batch
        enc.begin(W // 16, H // 16, P)                           # same handle, frame-by-frame pushes (staged path)
        for k in range(n):
            enc.push_frames(fr[k:k + 1])
        enc.sequence_stop()
        assert enc.drain()[0] == want, batch
        enc.begin(W // 16, H // 16, P)                           # 9 frames, then 14 more
        enc.push_frames(fr[:9]); enc.push_frames(fr[9:]); enc.sequence_stop()
        assert enc.drain()[0] == want, batch
        enc.close()
    want_p = ob.encode(fr, W // 16, H // 16, P, XL=6, YL=6, partial_px4=333)
    enc = pkg.Mpeg2Encoder(XL=6, YL=6)
    enc.set_limits(batch_frames=4, chunk_frames=4)
    assert enc.encode_sequence(fr, P, partial_px4=333) == want_p
    enc.close()
    import torch
    d = torch.from_numpy(fr[:20]).cuda()
    enc = pkg.Mpeg2Encoder(XL=6, YL=6)
    out = np.zeros(4 << 20, np.uint8)
    for chunk in (0, 4, 8):                                      # automatic (one chunk), 1 GOP, 2 GOPs per chunk
        enc.set_limits(chunk_frames=chunk)
        ln = enc.encode_gops_host(d.data_ptr(), 20, 0, W // 16, H // 16, P, out)
        assert pkg.finish_stream(pkg.sequence_header(W // 16, H // 16) + out[:ln].tobytes()) == ob.encode(fr[:20], W // 16, H // 16, P, XL=6, YL=6), chunk
    enc.close()


def test_streaming_full_size_multi_batch(pkg, ob, synth):
    """1920x1152 I+15P through the streaming C-ABI at the default thresholds: one GOP (106 MB) per batch, so 3 GOPs + 5
    frames cross the double-buffered H2D pipeline three times and end in the staged path - against the oracle, GOP by GOP"""
    import threading
    W, H, P, n = 1920, 1152, 15, 53
    fr = synth.s1_pan(9, n, W, H)
    enc = pkg.Mpeg2Encoder(XL=7, YL=7)
    got = enc.encode_sequence(fr, P)
    enc.close()
    bodies = [None] * 4
    def work(g):
        bodies[g] = ob.encode_range(fr[16 * g:16 * (g + 1)], 16 * g, W // 16, H // 16, P)
    th = [threading.Thread(target=work, args=(g,)) for g in range(4)]
    for t in th: t.start()
    for t in th: t.join()
    assert got == pkg.finish_stream(pkg.sequence_header(W // 16, H // 16) + b''.join(bodies))


def test_randomized_sequences(pkg, ob, synth):
    """a seeded sweep over sizes, lengths, GOP lengths, parameters, clip classes and stop positions (the reference's
    testbench exercises exactly one point of this space, TB:23-24,98-99,106)"""
    rng = np.random.default_rng(20261017)
    gens = [synth.s1_pan, synth.s2_white, synth.s3_dark, synth.s4_edges]
    encs = {}
    for case in range(28):
        W, H = 16 * int(rng.integers(4, 17)), 16 * int(rng.integers(4, 13))
        n = int(rng.integers(1, 8))
        P = int(rng.choice([0, 1, 2, 3, 5, 23, 255]))
        VL, Q = int(rng.integers(1, 4)), int(rng.integers(1, 5))
        gen = gens[int(rng.integers(0, 4))]
        partial = int(rng.integers(1, W * H // 4)) if rng.random() < 0.3 else 0
        fr = gen(1000 + case, n, W, H)
        if (VL, Q) not in encs:
            encs[(VL, Q)] = pkg.Mpeg2Encoder(XL=7, YL=7, VECTOR_LEVEL=VL, Q_LEVEL=Q)
        enc = encs[(VL, Q)]                                      # handles are reused across sequences, like the testbench's one instance
        if rng.random() < 0.5:
            enc.set_limits(batch_frames=int(rng.integers(1, 6)), chunk_frames=int(rng.integers(1, 6)))
        else:
            enc.set_limits(0, 0)
        got = enc.encode_sequence(fr, P, partial_px4=partial)
        want = ob.encode(fr, W // 16, H // 16, P, XL=7, YL=7, VL=VL, Q=Q, partial_px4=partial)
        assert got == want, (case, W, H, n, P, VL, Q, gen.__name__, partial)
    for e in encs.values():
        e.close()


def test_contract_errors_and_state_machine(pkg, synth):
    """error behaviour at the boundary: parameters outside README.md:79-84 are rejected, a new sequence cannot start
    while o_sequence_busy is high (README.md:222), pixels before begin / after stop are refused, a stop while idle is
    ignored (RTL:1090), and busy falls only when the o_last word has been pulled (RTL:1045-1047, 1095)"""
    M = pkg.M2VError
    for bad in (dict(XL=3), dict(XL=8), dict(YL=3), dict(VECTOR_LEVEL=0), dict(VECTOR_LEVEL=4), dict(Q_LEVEL=0), dict(Q_LEVEL=5)):
        with pytest.raises(M) as ei:
            pkg.Mpeg2Encoder(**bad)
        assert ei.value.code == -1                               # M2V_EINVAL
    enc = pkg.Mpeg2Encoder(XL=6, YL=6)
    fr = synth.s1_pan(1, 2, 64, 64)
    with pytest.raises(M) as ei:                                 # push before begin
        enc.push4(fr[0, 0, 0, :4], fr[0, 1, 0, :4], fr[0, 2, 0, :4])
    assert ei.value.code == -2                                   # M2V_ESTATE
    enc.sequence_stop()                                          # stop while idle: ignored
    assert not enc.sequence_busy and enc.pull() is None
    with pytest.raises(M):
        enc.begin(4, 4, 256)                                     # i_pframes_count is 8 bit
    assert enc.begin(2, 99, 1) == (4, 64)                        # clamp to [4, 2^YL] (RTL:985-991)
    enc.begin(4, 4, 1)                                           # not armed yet (no pixel): begin again is legal
    assert not enc.sequence_busy
    enc.push_frames(fr)
    assert enc.sequence_busy
    with pytest.raises(M) as ei:
        enc.begin(4, 4, 1)
    assert ei.value.code == -2
    enc.sequence_stop()
    with pytest.raises(M) as ei:                                 # pixels after the stop, before the stream is pulled
        enc.push_frames(fr)
    assert ei.value.code == -2
    words = []
    while True:
        assert enc.sequence_busy                                 # busy until the o_last word is out
        w = enc.pull()
        words.append(w[0])
        if w[1]:
            break
    assert not enc.sequence_busy and enc.pull() is None
    enc2 = pkg.Mpeg2Encoder(XL=6, YL=6)
    assert b''.join(words) == enc2.encode_sequence(fr, 1)
    assert enc.encode_sequence(fr, 1) == b''.join(words)          # the handle is reusable after the sequence ended
    enc.close(); enc2.close()


def test_long_stream_single_drain(pkg, ob, synth):
    """white noise at Q_LEVEL=1 codes > 1 byte per pixel: a stream of several MB pulled by ONE m2v_drain call (the
    multi-threaded queue -> caller copy) equals the oracle's"""
    W, H, n = 640, 480, 12
    fr = synth.s2_white(5, n, W, H)
    enc = pkg.Mpeg2Encoder(XL=6, YL=5, VECTOR_LEVEL=1, Q_LEVEL=1)
    enc.begin(W // 16, H // 16, 1)
    enc.push_frames(fr); enc.sequence_stop()
    buf = np.empty(64 << 20, np.uint8)
    k, last = enc.drain_into(buf)
    assert last and k > (4 << 20)
    assert buf[:k].tobytes() == ob.encode(fr, W // 16, H // 16, 1, XL=6, YL=5, VL=1, Q=1)
    enc.close()


def test_config1_testbench_clips(pkg):
    """BASELINE.json config 1 through the CUDA path: the testbench's three clips back to back on one instance with its default
    parameters (TB:23-24, 98-99, 106, 150).  The expected lengths and hashes were written by the reference RTL itself
    (tests/golden/clips_sha256.json); 1440x704 must come out at the 775 456 bytes the reference publishes (README.md:748).
    The clips are unpacked by `make -C oracle` into the git-ignored oracle/_ref/data and only exist where they travelled."""
    import hashlib, json
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    data = os.path.join(root, 'oracle', '_ref', 'data')
    meta = json.load(open(os.path.join(GOLD, 'clips_sha256.json')))
    if not all(os.path.exists(os.path.join(data, n + '.yuv')) for n in meta):
        pytest.skip('testbench clips not on this box')
    enc = pkg.Mpeg2Encoder(XL=7, YL=6, VECTOR_LEVEL=3, Q_LEVEL=2)
    for name in ('288x208', '640x320', '1440x704'):
        W, H = (int(x) for x in name.split('x'))
        raw = np.fromfile(os.path.join(data, name + '.yuv'), dtype=np.uint8)
        assert hashlib.sha256(raw.tobytes()).hexdigest() == meta[name]['input_sha256']
        fr = raw.reshape(-1, 3, H, W)
        assert fr.shape[0] == meta[name]['frames']
        got = enc.encode_sequence(fr, 23)
        assert len(got) == meta[name]['length'], name
        assert hashlib.sha256(got).hexdigest() == meta[name]['sha256'], name
    assert meta['1440x704']['length'] == 775456
    enc.close()


def test_testbench_replay_cli(pkg, ob, synth, tmp_path):
    """csrc/m2venc_tb.cpp = C++ host replaying TB:142-274 through the C-ABI: several videos back to back on one
    instance (TB:150), one frame per push like the testbench's frame loop, and the 4-pixel port (-push4)."""
    import subprocess
    exe = os.path.join(os.path.dirname(pkg.LIB_PATH), 'm2venc_tb')
    if not os.path.exists(exe):
        pytest.skip('m2venc_tb not built')
    vids = [(synth.s1_pan(1, 5, 96, 64), 96, 64), (synth.s4_edges(2, 3, 64, 80), 64, 80)]
    args = []
    for i, (fr, W, H) in enumerate(vids):
        fr.tofile(str(tmp_path / ('v%d.yuv' % i)))
        args += [str(tmp_path / ('v%d.yuv' % i)), str(W), str(H), str(tmp_path / ('v%d.m2v' % i))]
    for extra in ([], ['-push4']):
        subprocess.check_call([exe, '-XL', '6', '-YL', '6', '-VL', '3', '-Q', '2', '-P', '23'] + extra + args, stdout=subprocess.DEVNULL)
        for i, (fr, W, H) in enumerate(vids):
            assert open(str(tmp_path / ('v%d.m2v' % i)), 'rb').read() == ob.encode(fr, W // 16, H // 16, 23, XL=6, YL=6, VL=3, Q=2)

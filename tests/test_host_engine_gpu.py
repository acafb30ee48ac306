"""-m gpu tests of the host engine behind the C-ABI: several devices behind one handle (m2v_create_multi), several handles
in one process, the asynchronous chunk calls (m2v_gops_*), the regrow-and-rerun path of the body buffer, the pinned
helpers, the NCCL gather on real GPUs and the file-to-file CLI.  Bit-exact against the oracle; tests that need more than
one GPU skip on a single-GPU box."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch
    return torch.cuda.device_count()


def test_zero_frame_push_does_not_arm(pkg, ob, synth):
    """a push without pixels is no i_en: the sequence stays idle (RTL:1060-1065) and a stop is ignored (RTL:1090)"""
    enc = pkg.Mpeg2Encoder(XL=6, YL=6)
    enc.begin(4, 4, 1)
    enc.push_frames(np.zeros((0, 3, 64, 64), np.uint8))
    assert not enc.sequence_busy
    enc.sequence_stop()
    assert enc.pull() is None and not enc.sequence_busy
    fr = synth.s1_pan(3, 2, 64, 64)
    assert enc.encode_sequence(fr, 1) == ob.encode(fr, 4, 4, 1, XL=6, YL=6)
    enc.close()


def test_body_buffer_regrows_and_reruns(pkg, ob, synth):
    """white noise at Q_LEVEL=1 codes > 1 byte per pixel; with 1 byte per macroblock reserved every batch overflows its body
    buffer, is detected after the scans (nothing is written), regrown and run again - streaming and device-resident paths"""
    import torch
    W, H, P = 160, 96, 3
    fr = synth.s2_white(11, 13, W, H)
    want = ob.encode(fr, W // 16, H // 16, P, XL=6, YL=6, VL=2, Q=1)
    enc = pkg.Mpeg2Encoder(XL=6, YL=6, VECTOR_LEVEL=2, Q_LEVEL=1)
    enc.set_body_reserve(1)
    enc.set_limits(batch_frames=4)
    assert enc.encode_sequence(fr, P) == want
    enc.set_body_reserve(1)
    d = torch.from_numpy(fr[:12]).cuda()
    out = np.zeros(8 << 20, np.uint8)
    ln = enc.encode_gops_host(d.data_ptr(), 12, 0, W // 16, H // 16, P, out)
    assert pkg.finish_stream(pkg.sequence_header(W // 16, H // 16) + out[:ln].tobytes()) == ob.encode(fr[:12], W // 16, H // 16, P, XL=6, YL=6, VL=2, Q=1)
    enc.close()


def test_async_chunks_two_slots(pkg, ob, synth):
    """m2v_gops_submit / size / fetch / wait: four chunks through the two slots, bodies fetched into pinned memory while the
    next chunk is encoded; equal to the synchronous call and to the oracle"""
    import torch
    W, H, P, n = 128, 96, 3, 32
    fr = synth.s1_pan(21, n, W, H)
    d = torch.from_numpy(fr).cuda()
    torch.cuda.synchronize()
    enc = pkg.Mpeg2Encoder(XL=6, YL=6)
    fsz = 3 * W * H
    pin = pkg.PinnedArray(4 << 20)
    chunks = [(0, 8), (8, 8), (16, 12), (28, 4)]
    sizes, off = [], 0
    enc.gops_submit(d.data_ptr(), chunks[0][1], 0, W // 16, H // 16, P, 0)
    for i, (f0, nf) in enumerate(chunks):
        if i + 1 < len(chunks):
            g0, gn = chunks[i + 1]
            enc.gops_submit(d.data_ptr() + g0 * fsz, gn, g0, W // 16, H // 16, P, (i + 1) & 1)
        k = enc.gops_size(i & 1)
        enc.gops_fetch(i & 1, pin.ptr + off, k)
        sizes.append(k); off += k
    enc.gops_wait(0); enc.gops_wait(1)
    got = pin.array[:off].tobytes()
    out = np.zeros(4 << 20, np.uint8)
    ln = enc.encode_gops_host(d.data_ptr(), n, 0, W // 16, H // 16, P, out)
    assert got == out[:ln].tobytes()
    assert pkg.finish_stream(pkg.sequence_header(W // 16, H // 16) + got) == ob.encode(fr, W // 16, H // 16, P, XL=6, YL=6)
    pin.close(); enc.close()


def test_pinned_source_and_pageable_source_agree(pkg, synth):
    fr = synth.s1_pan(8, 10, 96, 64)
    pin = pkg.PinnedArray(fr.nbytes)
    pin.array[:] = fr.reshape(-1)
    enc = pkg.Mpeg2Encoder(XL=6, YL=6)
    a = enc.encode_sequence(fr, 4)
    b = enc.encode_sequence(pin.array.reshape(fr.shape), 4)
    assert a == b
    pin.close(); enc.close()


def test_multi_entry_point_with_one_device(pkg, ob, synth):
    """m2v_create_multi(1, ...) is the same instance as m2v_create(...): same engine, same stream; out-of-range device counts are
    refused (more devices than the box has: M2V_ENODEV, not a crash)"""
    import torch
    fr = synth.s4_edges(33, 11, 128, 80)
    want = ob.encode(fr, 8, 5, 4, XL=6, YL=6)
    enc = pkg.Mpeg2Encoder(XL=6, YL=6, ndev=1, force_multi=True)
    assert enc.device_count == 1
    assert enc.encode_sequence(fr, 4) == want
    enc.close()
    n = torch.cuda.device_count()
    for bad, code in [(0, pkg.M2V_EINVAL), (9, pkg.M2V_EINVAL)] + ([(n + 1, pkg.M2V_ENODEV)] if n < 8 else []):
        with pytest.raises(pkg.M2VError) as ei:
            pkg.Mpeg2Encoder(XL=6, YL=6, ndev=bad, force_multi=True)
        assert ei.value.code == code, bad


def test_macroblocks_longer_than_the_cached_slot(pkg, ob, synth):
    """K2's count pass caches 1024 bits of a macroblock's bitstring for the write pass; white noise at Q_LEVEL=1 codes well
    over 2000 bits per macroblock, so every macroblock takes the second walk - beside short ones in the same warp (a dark,
    flat clip appended: a few dozen bits per macroblock)"""
    W, H, P = 128, 96, 2
    fr = np.concatenate([synth.s2_white(7, 4, W, H), synth.s3_dark(8, 5, W, H), synth.s2_white(9, 3, W, H)])
    enc = pkg.Mpeg2Encoder(XL=6, YL=6, VECTOR_LEVEL=1, Q_LEVEL=1)
    got = enc.encode_sequence(fr, P)
    enc.close()
    assert got == ob.encode(fr, W // 16, H // 16, P, XL=6, YL=6, VL=1, Q=1)
    assert len(got) * 8 / (len(fr) * (W // 16) * (H // 16)) > 1024


def test_registered_host_memory_as_source_and_sink(pkg, ob, synth):
    """m2v_register_host pins memory the caller owns (here numpy arrays): frames pushed from it, words drained into it"""
    fr = np.ascontiguousarray(synth.s1_pan(12, 9, 160, 96))
    sink = np.zeros(4 << 20, np.uint8)
    L = pkg.lib()
    assert L.m2v_register_host(fr.ctypes.data, fr.nbytes) == 0 and L.m2v_register_host(sink.ctypes.data, sink.nbytes) == 0
    enc = pkg.Mpeg2Encoder(XL=6, YL=6)
    enc.begin(10, 6, 3)
    enc.push_frames(fr); enc.sequence_stop()
    n, last = enc.drain_into(sink)
    assert last and sink[:n].tobytes() == ob.encode(fr, 10, 6, 3, XL=6, YL=6)
    enc.close()
    assert L.m2v_unregister_host(fr.ctypes.data) == 0 and L.m2v_unregister_host(sink.ctypes.data) == 0
    assert L.m2v_register_host(0, 4096) != 0                       # refused, reported - and the handle after it works
    enc = pkg.Mpeg2Encoder(XL=6, YL=6)
    assert enc.encode_sequence(fr[:3], 1) == ob.encode(fr[:3], 10, 6, 1, XL=6, YL=6)
    enc.close()


def test_two_handles_on_two_devices_in_one_process(pkg, ob, synth):
    """the K1 launch configuration (dynamic shared memory attribute, persistent grid size) is per device: a second handle on
    another GPU of the same process must run P-frames too"""
    import torch
    if _ngpu() < 2:
        pytest.skip('needs 2 GPUs')
    fr = synth.s1_pan(41, 9, 160, 96)
    want = ob.encode(fr, 10, 6, 3, XL=6, YL=6)
    encs = []
    for dev in (0, 1):
        torch.cuda.set_device(dev)
        encs.append(pkg.Mpeg2Encoder(XL=6, YL=6))
    torch.cuda.set_device(0)
    for enc in encs + encs[::-1]:
        assert enc.encode_sequence(fr, 3) == want
    for enc in encs:
        enc.close()


@pytest.mark.parametrize('ndev', [2, 4, 8])
def test_one_instance_on_n_devices(pkg, ob, synth, ndev):
    """m2v_create_multi: the streaming calls deal whole-GOP batches to the devices; the ordered stream equals the single-device
    one and the oracle's (bulk push, frame-by-frame pushes, a partial last frame, tiny forced batches, sequences back to back)"""
    if _ngpu() < ndev:
        pytest.skip('needs %d GPUs' % ndev)
    W, H, P, n = 160, 96, 3, 45
    fr = synth.s1_pan(51, n, W, H)
    want = ob.encode(fr, W // 16, H // 16, P, XL=6, YL=6)
    one = pkg.Mpeg2Encoder(XL=6, YL=6)
    multi = pkg.Mpeg2Encoder(XL=6, YL=6, ndev=ndev)
    assert multi.device_count == ndev
    assert one.encode_sequence(fr, P) == want
    assert multi.encode_sequence(fr, P) == want
    multi.set_limits(batch_frames=4)
    assert multi.encode_sequence(fr, P) == want
    multi.begin(W // 16, H // 16, P)
    for k in range(n):
        multi.push_frames(fr[k:k + 1])
    multi.sequence_stop()
    assert multi.drain()[0] == want
    assert multi.encode_sequence(fr, P, partial_px4=77) == ob.encode(fr, W // 16, H // 16, P, XL=6, YL=6, partial_px4=77)
    multi.set_limits(0, 0)
    W2, H2 = 1920, 1152
    fr2 = synth.s1_pan(52, 40, W2, H2)
    big = pkg.Mpeg2Encoder(XL=7, YL=7, ndev=ndev)
    ref = pkg.Mpeg2Encoder(XL=7, YL=7)
    assert big.encode_sequence(fr2, 15) == ref.encode_sequence(fr2, 15)
    for e in (one, multi, big, ref):
        e.close()


def test_nccl_gather_two_ranks():
    """sharding.gather_bodies over NCCL on two GPUs: two ranks encode their GOP blocks, rank 0 assembles and compares with
    the oracle (tests/nccl_gather_worker.py, launched with torch.distributed.run)"""
    if _ngpu() < 2:
        pytest.skip('needs 2 GPUs')
    env = dict(os.environ)
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
                        '--master-port', '29533', os.path.join(ROOT, 'tests', 'nccl_gather_worker.py')], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert 'NCCL GATHER OK' in r.stdout and 'HOST ARENA OK' in r.stdout


def test_testbench_cli_chunked_and_multi_gpu(pkg, ob, synth, tmp_path):
    """m2venc_tb: the read-ahead / write-behind file->file path (whole chunks pushed from pinned buffers), the frame-by-frame
    mode, and -gpus 2 where two GPUs exist"""
    exe = os.path.join(os.path.dirname(pkg.LIB_PATH), 'm2venc_tb')
    if not os.path.exists(exe):
        pytest.skip('m2venc_tb not built')
    fr = synth.s1_pan(61, 21, 160, 96)
    fr.tofile(str(tmp_path / 'v.yuv'))
    want = ob.encode(fr, 10, 6, 3, XL=6, YL=6, VL=3, Q=2)
    modes = [[], ['-chunk', '8'], ['-frame'], ['-chunk', '4', '-readers', '3']]
    if _ngpu() >= 2:
        modes.append(['-gpus', '2', '-chunk', '8'])
    for extra in modes:
        out = str(tmp_path / 'v.m2v')
        subprocess.check_call([exe, '-XL', '6', '-YL', '6', '-P', '3'] + extra + [str(tmp_path / 'v.yuv'), '160', '96', out], stdout=subprocess.DEVNULL)
        assert open(out, 'rb').read() == want, extra

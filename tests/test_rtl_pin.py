"""Pins the oracle against THE REFERENCE ITSELF: /root/reference/RTL/mpeg2encoder.v translated to C++ by
oracle/vl2c.py and driven clock by clock by a replay of the reference testbench (oracle/rtl_tb.cpp).
Where the reference sources are present (build container) the model is rebuilt on demand; on the GPU box
the prebuilt oracle/_ref/*.so that travelled with the snapshot is used; with neither the tests skip."""
import hashlib
import json
import os
import zipfile

import numpy as np
import pytest

import rtl_ref_binding as rb

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, 'golden')
REF_ZIP = '/root/reference/SIM/data.zip'
HAVE_SRC = os.path.exists(rb.RTL)


def need(XL=7, YL=6, VL=3, Q=2):
    if not rb.available(XL, YL, VL, Q):
        pytest.skip('no reference sources and no prebuilt oracle/_ref model for these parameters')


def test_testbench_default_run_matches_published_size_and_oracle(ob):
    """config 1: the reference's own clips through the reference's own logic with the testbench defaults
    (XL=7,YL=6,VECTOR_LEVEL=3,Q_LEVEL=2,i_pframes_count=23; TB:23-24,98-99,106), back to back on one instance
    (TB:150).  1440x704.yuv must give the 775 456 bytes the README publishes (README.md:748), and every stream
    must equal the oracle's byte for byte."""
    if not os.path.exists(REF_ZIP):
        pytest.skip('reference clips not present on this box')
    need()
    meta = json.load(open(os.path.join(GOLD, 'clips_sha256.json')))
    z = zipfile.ZipFile(REF_ZIP)
    r = rb.RtlRef(7, 6, 3, 2)
    clips = [('288x208', 288, 208), ('1440x704', 1440, 704)]
    if os.environ.get('M2V_FULL_TB'):
        clips.insert(1, ('640x320', 640, 320))
    for name, W, H in clips:
        fr = np.frombuffer(z.read('data/%s.yuv' % name), dtype=np.uint8).reshape(-1, 3, H, W)
        got = r.sequence(fr, W // 16, H // 16, 23)
        assert len(got) == meta[name]['length'], name
        assert hashlib.sha256(got).hexdigest() == meta[name]['sha256'], name          # == oracle (make_golden.py)
        if name == '1440x704':
            assert len(got) == 775456
            assert got == ob.encode(fr, W // 16, H // 16, 23, XL=7, YL=6, VL=3, Q=2)


@pytest.mark.parametrize('VL', [1, 2, 3])
@pytest.mark.parametrize('Q', [1, 2, 3, 4])
def test_parameter_grid_rtl_equals_oracle(ob, synth, VL, Q):
    if not HAVE_SRC:
        need(6, 6, VL, Q)
    r = rb.RtlRef(6, 6, VL, Q)
    seed = 10 * VL + Q
    clips = [(synth.s1_pan(seed, 4, 96, 64), 2), (synth.s4_edges(seed, 3, 64, 80), 7), (synth.s2_white(seed, 2, 64, 64), 1),
             (synth.s3_dark(seed, 4, 80, 64), 0)]
    for fr, P in clips:                                           # back to back on one instance, no reset in between
        _, _, H, W = fr.shape
        assert r.sequence(fr, W // 16, H // 16, P) == ob.encode(fr, W // 16, H // 16, P, XL=6, YL=6, VL=VL, Q=Q)


def test_stop_mid_frame_clamp_and_bubbles(ob, synth):
    need(7, 6, 3, 2)
    r = rb.RtlRef(7, 6, 3, 2)
    fr = synth.s1_pan(3, 3, 80, 64)
    # i_sequence_stop in the middle of the third frame: the rest is padded with black (RTL:1036-1056)
    for px4 in (1, 333, 80 * 64 // 4 - 1):
        assert r.sequence(fr, 5, 4, 5, partial_px4=px4) == ob.encode(fr, 5, 4, 5, XL=7, YL=6, partial_px4=px4)
    # input bubbles (TB:233) do not change the stream
    assert r.sequence(fr, 5, 4, 5, bubble_seed=12345) == ob.encode(fr, 5, 4, 5, XL=7, YL=6)
    # size clamp (RTL:985-991): i_xsize16=2 -> 4, i_ysize16=100 -> 64 (YL=6)
    tall = synth.s1_pan(4, 1, 64, 1024)
    assert r.sequence(tall, 2, 100, 0) == ob.encode(tall, 2, 100, 0, XL=7, YL=6)
    assert r.sequence(tall, 4, 64, 0) == ob.encode(tall, 4, 64, 0, XL=7, YL=6)


def test_long_gop_time_code_and_pframe_wrap(ob):
    need(7, 6, 3, 2)
    r = rb.RtlRef(7, 6, 3, 2)
    rng = np.random.default_rng(5)
    base = rng.integers(0, 256, (1, 3, 64, 64), dtype=np.uint8)
    fr = np.repeat(base, 60, axis=0)
    fr[:, 0, :8, :8] = rng.integers(0, 256, (60, 8, 8), dtype=np.uint8)
    for P in (255, 23, 4):
        assert r.sequence(fr, 4, 4, P) == ob.encode(fr, 4, 4, P, XL=7, YL=6)


def test_committed_fixtures_are_rtl_outputs(ob):
    """tests/golden/*.m2v were written by the RTL model (make_golden.py); the oracle must reproduce them and,
    where the model is available, so must a fresh RTL run."""
    meta = json.load(open(os.path.join(GOLD, 'rtl_fixtures.json')))
    for name, m in meta.items():
        fr = np.fromfile(os.path.join(GOLD, name + '.yuv'), dtype=np.uint8).reshape(m['frames'], 3, m['H'], m['W'])
        want = open(os.path.join(GOLD, name + '.m2v'), 'rb').read()
        assert hashlib.sha256(want).hexdigest() == m['sha256']
        kw = dict(XL=m['XL'], YL=m['YL'], VL=m['VL'], Q=m['Q'], partial_px4=m['partial_px4'])
        assert ob.encode(fr, m['W'] // 16, m['H'] // 16, m['P'], **kw) == want
        if rb.available(m['XL'], m['YL'], m['VL'], m['Q']):
            r = rb.RtlRef(m['XL'], m['YL'], m['VL'], m['Q'])
            assert r.sequence(fr, m['W'] // 16, m['H'] // 16, m['P'], partial_px4=m['partial_px4']) == want

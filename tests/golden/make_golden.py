#!/usr/bin/env python3
"""Regenerates tests/golden/*.  Runs only in the build container (needs /root/reference).

Every stream here is written by THE REFERENCE RTL ITSELF - /root/reference/RTL/mpeg2encoder.v translated to
C++ by oracle/vl2c.py and driven by the testbench replay oracle/rtl_tb.cpp (tests/rtl_ref_binding.py) - and is
checked to be byte-identical to the oracle's output before it is written.

  clipA_64x64.yuv/.m2v : top-left 64x64 crop of the first 5 frames of SIM/data.zip:288x208.yuv, testbench
                         parameters (XL=7,YL=6,VECTOR_LEVEL=3,Q_LEVEL=2,i_pframes_count=23; TB:23-24,98-99,106)
  clipB_64x64.*        : S2 white noise, 3 frames, VECTOR_LEVEL=1, Q_LEVEL=1, i_pframes_count=1 (escape codes, intra in P)
  clipC_96x64.*        : S4 edges, 3 frames + 500 pixel groups of a 4th, VECTOR_LEVEL=2, Q_LEVEL=4, i_pframes_count=2
  rtl_fixtures.json    : parameters + sha256 of the above
  clips_sha256.json    : sha256 + length of the RTL's streams for the three bundled clips with the testbench
                         defaults.  1440x704 is 775456 bytes - the number the reference publishes (README.md:748).
"""
import hashlib, json, os, sys, zipfile
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle_binding as ob
import rtl_ref_binding as rb
import __graft_entry__ as ge
synth = ge.load_synth()

z = zipfile.ZipFile('/root/reference/SIM/data.zip')
clips = {'288x208': (288, 208), '640x320': (640, 320), '1440x704': (1440, 704)}
meta = {}
tb = rb.RtlRef(7, 6, 3, 2)
crop = None
for name, (W, H) in clips.items():                                  # back to back on one instance (TB:150)
    raw = np.frombuffer(z.read('data/%s.yuv' % name), dtype=np.uint8)
    n = raw.size // (W * H * 3)
    fr = raw.reshape(n, 3, H, W)
    out = tb.sequence(fr, W // 16, H // 16, 23)
    assert out == ob.encode(fr, W // 16, H // 16, 23, XL=7, YL=6, VL=3, Q=2), name
    meta[name] = dict(frames=n, length=len(out), sha256=hashlib.sha256(out).hexdigest(),
                      input_sha256=hashlib.sha256(raw.tobytes()).hexdigest(), producer='reference RTL via oracle/vl2c.py')
    if name == '288x208':
        crop = np.ascontiguousarray(fr[:5, :, :64, :64])
assert meta['1440x704']['length'] == 775456
json.dump(meta, open(os.path.join(HERE, 'clips_sha256.json'), 'w'), indent=1)

fixtures = {
    'clipA_64x64': dict(frames=crop, XL=7, YL=6, VL=3, Q=2, P=23, partial_px4=0),
    'clipB_64x64': dict(frames=synth.s2_white(11, 3, 64, 64), XL=6, YL=6, VL=1, Q=1, P=1, partial_px4=0),
    'clipC_96x64': dict(frames=synth.s4_edges(12, 4, 96, 64), XL=6, YL=6, VL=2, Q=4, P=2, partial_px4=500),
}
fmeta = {}
for name, f in fixtures.items():
    fr = f['frames']; n, _, H, W = fr.shape
    r = rb.RtlRef(f['XL'], f['YL'], f['VL'], f['Q'])
    out = r.sequence(fr, W // 16, H // 16, f['P'], partial_px4=f['partial_px4'])
    assert out == ob.encode(fr, W // 16, H // 16, f['P'], XL=f['XL'], YL=f['YL'], VL=f['VL'], Q=f['Q'], partial_px4=f['partial_px4']), name
    fr.tofile(os.path.join(HERE, name + '.yuv'))
    open(os.path.join(HERE, name + '.m2v'), 'wb').write(out)
    fmeta[name] = dict(frames=n, W=W, H=H, XL=f['XL'], YL=f['YL'], VL=f['VL'], Q=f['Q'], P=f['P'], partial_px4=f['partial_px4'],
                       length=len(out), sha256=hashlib.sha256(out).hexdigest())
json.dump(fmeta, open(os.path.join(HERE, 'rtl_fixtures.json'), 'w'), indent=1)
print(json.dumps(meta, indent=1)); print(json.dumps(fmeta, indent=1))

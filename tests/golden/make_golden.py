#!/usr/bin/env python3
"""Regenerates tests/golden/*.  Runs only in the build container (needs /root/reference).

  clipA_64x64.yuv / .m2v   : top-left 64x64 crop of the first 5 frames of SIM/data.zip:288x208.yuv,
                             encoded by the oracle with the testbench parameters (VECTOR_LEVEL=3,
                             Q_LEVEL=2, i_pframes_count=23 -> I+4P here) [TB:98-99,106]
  clips_sha256.json        : sha256 + length of the oracle's streams for the three bundled clips with
                             the testbench defaults (XL=7,YL=6; TB:23-24).  1440x704 must be 775456
                             bytes (README.md:748) - the only number the reference publishes.
There is no simulator in the image, so these are ORACLE outputs (regression pins), not RTL outputs.
"""
import hashlib, json, os, sys, zipfile
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_binding as ob

z = zipfile.ZipFile('/root/reference/SIM/data.zip')
clips = {'288x208': (288, 208), '640x320': (640, 320), '1440x704': (1440, 704)}
meta = {}
for name, (W, H) in clips.items():
    raw = np.frombuffer(z.read('data/%s.yuv' % name), dtype=np.uint8)
    n = raw.size // (W * H * 3)
    fr = raw.reshape(n, 3, H, W)
    out = ob.encode(fr, W // 16, H // 16, 23, XL=7, YL=6, VL=3, Q=2)
    meta[name] = dict(frames=n, length=len(out), sha256=hashlib.sha256(out).hexdigest(),
                      input_sha256=hashlib.sha256(raw.tobytes()).hexdigest())
    if name == '288x208':
        crop = np.ascontiguousarray(fr[:5, :, :64, :64])
        crop.tofile(os.path.join(HERE, 'clipA_64x64.yuv'))
        open(os.path.join(HERE, 'clipA_64x64.m2v'), 'wb').write(ob.encode(crop, 4, 4, 23, XL=7, YL=6, VL=3, Q=2))
assert meta['1440x704']['length'] == 775456
json.dump(meta, open(os.path.join(HERE, 'clips_sha256.json'), 'w'), indent=1)
print(json.dumps(meta, indent=1))

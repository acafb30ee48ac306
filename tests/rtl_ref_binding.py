"""ctypes binding of oracle/_ref/librtl_ref_*.so: the reference RTL itself (translated to C++ by
oracle/vl2c.py, driven by the testbench replay oracle/rtl_tb.cpp).  TEST INFRASTRUCTURE - only tests/ and
bench.py's CPU-baseline legs may use it.  In the build container the library is (re)built on demand from
/root/reference; elsewhere (GPU box) only prebuilt files that travelled with the snapshot are used."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RTL = '/root/reference/RTL/mpeg2encoder.v'
_libs = {}


def available(XL=7, YL=6, VL=3, Q=2):
    return os.path.exists(RTL) or os.path.exists(_path(XL, YL, VL, Q))


def _path(XL, YL, VL, Q):
    return os.path.join(ROOT, 'oracle', '_ref', 'librtl_ref_XL%d_YL%d_VL%d_Q%d.so' % (XL, YL, VL, Q))


def lib(XL=7, YL=6, VL=3, Q=2):
    key = (XL, YL, VL, Q)
    if key not in _libs:
        if os.path.exists(RTL):
            subprocess.check_call(['make', '-s', '-C', os.path.join(ROOT, 'oracle'), 'ref', 'XL=%d' % XL, 'YL=%d' % YL, 'VL=%d' % VL, 'Q=%d' % Q])
        L = C.CDLL(_path(*key))
        L.rtl_ref_create.restype = C.c_void_p
        L.rtl_ref_destroy.argtypes = [C.c_void_p]
        L.rtl_ref_clocks.argtypes = [C.c_void_p]; L.rtl_ref_clocks.restype = C.c_long
        L.rtl_ref_sequence.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_long, C.c_long, C.c_uint, C.c_void_p,
                                       C.c_size_t, C.POINTER(C.c_size_t)]
        _libs[key] = L
    return _libs[key]


class RtlRef:
    """one module instance after reset; sequence() = one video, as the testbench drives it"""

    def __init__(self, XL=7, YL=6, VL=3, Q=2):
        self.L = lib(XL, YL, VL, Q)
        self.XL, self.YL = XL, YL
        self.h = C.c_void_p(self.L.rtl_ref_create())

    def sequence(self, frames, xsize16, ysize16, pframes, partial_px4=0, bubble_seed=0):
        f = np.ascontiguousarray(frames, dtype=np.uint8)
        n = f.shape[0]
        cap = 4096 + n * f[0].size * 2
        out = np.zeros(cap, np.uint8)
        ln = C.c_size_t(0)
        rc = self.L.rtl_ref_sequence(self.h, f.ctypes.data, xsize16, ysize16, pframes, n - (1 if partial_px4 else 0), partial_px4,
                                     bubble_seed, out.ctypes.data, cap, C.byref(ln))
        if rc:
            raise RuntimeError('rtl_ref_sequence rc=%d' % rc)
        return out[:ln.value].tobytes()

    @property
    def clocks(self):
        return self.L.rtl_ref_clocks(self.h)

    def close(self):
        if self.h:
            self.L.rtl_ref_destroy(self.h); self.h = None

    __del__ = close

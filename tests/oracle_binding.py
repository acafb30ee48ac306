"""ctypes binding of oracle/build/libm2v_oracle.so (TEST INFRASTRUCTURE; see oracle/m2v_oracle.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
import ctypes as C
import os
import subprocess
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB = None


class Dbg(C.Structure):
    _fields_ = [('mb_inter', C.c_void_p), ('mb_mvx', C.c_void_p), ('mb_mvy', C.c_void_p),
                ('mb_cbp', C.c_void_p), ('coefs', C.c_void_p), ('recon', C.c_void_p)]


def build():
    subprocess.check_call(['make', '-s', '-C', os.path.join(ROOT, 'oracle')])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(ROOT, 'oracle', 'build', 'libm2v_oracle.so')
        src = os.path.join(ROOT, 'oracle', 'm2v_oracle.c')
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
            build()
        L = C.CDLL(path)
        L.m2v_oracle_encode.restype = C.c_int
        L.m2v_oracle_encode.argtypes = [C.c_int] * 7 + [C.c_void_p, C.c_long, C.c_long, C.c_void_p,
                                                        C.c_size_t, C.POINTER(C.c_size_t), C.c_void_p]
        L.m2v_oracle_encode_range.restype = C.c_int
        L.m2v_oracle_encode_range.argtypes = [C.c_int] * 5 + [C.c_void_p, C.c_long, C.c_long, C.c_void_p,
                                                              C.c_size_t, C.POINTER(C.c_size_t), C.c_void_p]
        L.m2v_oracle_tail_len.restype = C.c_size_t
        L.m2v_oracle_tail_len.argtypes = [C.c_size_t]
        L.m2v_oracle_find_min10.argtypes = [C.POINTER(C.c_int)]
        L.m2v_oracle_put_ac.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_uint32)]
        _LIB = L
    return _LIB


def clamp16(s, L):
    return lib().m2v_oracle_clamp16(int(s), int(L))


def encode(frames, xsize16, ysize16, pframes, XL=7, YL=7, VL=3, Q=2, partial_px4=0, want_dbg=False):
    """frames: uint8 array [n, 3, H, W] (planar yuv444p, clamped geometry).  If partial_px4>0 the
    LAST frame in `frames` is the partial one.  Returns bytes (and a dict of dumps if want_dbg)."""
    L = lib()
    frames = np.ascontiguousarray(frames, dtype=np.uint8)
    n = frames.shape[0]
    nfull = n - (1 if partial_px4 > 0 else 0)
    mbw, mbh = clamp16(xsize16, XL), clamp16(ysize16, YL)
    W, H = mbw * 16, mbh * 16
    assert frames.shape[1:] == (3, H, W), (frames.shape, H, W)
    cap = 64 + n * (64 + mbh * 8 + mbw * mbh * 1200) + 64
    out = np.zeros(cap, dtype=np.uint8)
    outlen = C.c_size_t(0)
    dbg = None
    dumps = {}
    if want_dbg:
        nmb = n * mbw * mbh
        dumps = dict(mb_inter=np.zeros(nmb, np.int8), mb_mvx=np.zeros(nmb, np.int8),
                     mb_mvy=np.zeros(nmb, np.int8), mb_cbp=np.zeros(nmb, np.uint8),
                     coefs=np.zeros((nmb, 6, 64), np.int16), recon=np.zeros((n, W * H * 3 // 2), np.uint8))
        dbg = Dbg(*[dumps[k].ctypes.data for k in ('mb_inter', 'mb_mvx', 'mb_mvy', 'mb_cbp', 'coefs', 'recon')])
    rc = L.m2v_oracle_encode(XL, YL, VL, Q, xsize16, ysize16, pframes, frames.ctypes.data, nfull, partial_px4,
                             out.ctypes.data, cap, C.byref(outlen), C.byref(dbg) if dbg else None)
    if rc != 0:
        raise RuntimeError('oracle encode failed rc=%d' % rc)
    data = out[:outlen.value].tobytes()
    return (data, dumps) if want_dbg else data


def encode_range(frames, n0, mbw, mbh, pframes, VL=3, Q=2):
    L = lib()
    frames = np.ascontiguousarray(frames, dtype=np.uint8)
    n = frames.shape[0]
    cap = 64 + n * (64 + mbh * 8 + mbw * mbh * 1200)
    out = np.zeros(cap, dtype=np.uint8)
    outlen = C.c_size_t(0)
    rc = L.m2v_oracle_encode_range(VL, Q, mbw, mbh, pframes, frames.ctypes.data, n0, n0 + n,
                                   out.ctypes.data, cap, C.byref(outlen), None)
    if rc != 0:
        raise RuntimeError('oracle encode_range failed rc=%d' % rc)
    return out[:outlen.value].tobytes()


def seq_header(mbw, mbh):
    b = (C.c_uint8 * 34)()
    lib().m2v_oracle_seq_header(mbw, mbh, b)
    return bytes(b)

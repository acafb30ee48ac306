"""fpga-mpeg2-encoder_b200 - B200-native MPEG-2 I/P encoder behind the reference's streaming contract.

Host-side mirror of the module interface of the reference core (RTL/mpeg2encoder.v:10-38): the
static parameters XL/YL/VECTOR_LEVEL/Q_LEVEL, the per-sequence configuration
i_xsize16/i_ysize16/i_pframes_count, the 4-pixel-per-cycle YUV 4:4:4 input, i_sequence_stop /
o_sequence_busy and the 32-byte-word output.  Everything is a thin ctypes layer over the C-ABI in
include/m2venc.h (libm2venc.so, hand-written sm_100a CUDA).  There is NO CPU fallback: importing
works anywhere (the library only needs libcudart symbols that are linked statically), but creating
an encoder without a B200 raises.

The directory name contains '-', so load it with __graft_entry__.load_package() (importlib by
path); the module registers itself as `fpga_mpeg2_encoder_b200`.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libm2venc.so')

M2V_OK, M2V_EINVAL, M2V_ESTATE, M2V_ENODEV, M2V_ENOMEM, M2V_ECUDA, M2V_ESPACE = 0, -1, -2, -3, -4, -5, -6
_ERRNAMES = {-1: 'EINVAL', -2: 'ESTATE', -3: 'ENODEV', -4: 'ENOMEM', -5: 'ECUDA', -6: 'ESPACE'}

# every symbol include/m2venc.h declares (tests check the export list against the header)
ABI_SYMBOLS = [
    'm2v_create', 'm2v_destroy', 'm2v_last_error', 'm2v_begin', 'm2v_push4', 'm2v_push_frames', 'm2v_stop',
    'm2v_busy', 'm2v_pull', 'm2v_drain', 'm2v_encode_gops_device', 'm2v_encode_gops_host',
    'm2v_sequence_header', 'm2v_finish_stream', 'm2v_debug_copy', 'm2v_launch_count', 'm2v_kernel_ms',
    'm2v_set_timing', 'm2v_set_limits', 'm2v_set_body_reserve', 'm2v_create_multi', 'm2v_device_count',
    'm2v_gops_submit', 'm2v_gops_size', 'm2v_gops_fetch', 'm2v_gops_wait', 'm2v_gops_body',
    'm2v_alloc_host', 'm2v_free_host', 'm2v_register_host', 'm2v_unregister_host',
]


class M2VError(RuntimeError):
    def __init__(self, code, msg=''):
        super().__init__('m2venc error %s (%d) %s' % (_ERRNAMES.get(code, '?'), code, msg))
        self.code = code


def build_library(verbose=False):
    """Compile libm2venc.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    subprocess.check_call(['make', '-C', os.path.join(_HERE, 'csrc'), os.path.join('..', 'libm2venc.so')] +
                          ([] if verbose else ['-s']))


_lib = None


def lib():
    """Load the C-ABI library.  Fails loudly when it has not been built - no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError('libm2venc.so is missing (%s): run __graft_entry__.build()' % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        vp, ip, sz = C.c_void_p, C.c_int, C.c_size_t
        L.m2v_create.argtypes = [ip, ip, ip, ip, C.POINTER(vp)]
        L.m2v_destroy.argtypes = [vp]; L.m2v_destroy.restype = None
        L.m2v_last_error.argtypes = [vp]; L.m2v_last_error.restype = C.c_char_p
        L.m2v_begin.argtypes = [vp, ip, ip, ip, C.POINTER(ip), C.POINTER(ip)]
        L.m2v_push4.argtypes = [vp, vp, vp, vp]
        L.m2v_push_frames.argtypes = [vp, vp, C.c_long]
        L.m2v_stop.argtypes = [vp]
        L.m2v_busy.argtypes = [vp]
        L.m2v_pull.argtypes = [vp, vp, C.POINTER(ip)]
        L.m2v_drain.argtypes = [vp, vp, sz, C.POINTER(sz), C.POINTER(ip)]
        L.m2v_encode_gops_device.argtypes = [vp, ip, ip, ip, vp, C.c_long, C.c_long, C.POINTER(vp), C.POINTER(sz)]
        L.m2v_encode_gops_host.argtypes = [vp, ip, ip, ip, vp, C.c_long, C.c_long, vp, sz, C.POINTER(sz)]
        L.m2v_sequence_header.argtypes = [ip, ip, vp]
        L.m2v_finish_stream.argtypes = [vp, sz, sz, C.POINTER(sz)]
        L.m2v_debug_copy.argtypes = [vp, vp, vp, C.c_long]
        L.m2v_launch_count.argtypes = [vp]; L.m2v_launch_count.restype = C.c_long
        L.m2v_kernel_ms.argtypes = [vp, vp]
        L.m2v_set_timing.argtypes = [vp, ip]
        L.m2v_set_limits.argtypes = [vp, C.c_long, C.c_long]
        L.m2v_set_body_reserve.argtypes = [vp, C.c_long]
        L.m2v_create_multi.argtypes = [ip, ip, ip, ip, ip, C.POINTER(vp)]
        L.m2v_device_count.argtypes = [vp]
        L.m2v_gops_submit.argtypes = [vp, ip, ip, ip, vp, C.c_long, C.c_long, ip]
        L.m2v_gops_size.argtypes = [vp, ip, C.POINTER(sz)]
        L.m2v_gops_fetch.argtypes = [vp, ip, vp, sz]
        L.m2v_gops_wait.argtypes = [vp, ip]
        L.m2v_gops_body.argtypes = [vp, ip, C.POINTER(vp)]
        L.m2v_alloc_host.argtypes = [sz]; L.m2v_alloc_host.restype = vp
        L.m2v_free_host.argtypes = [vp]; L.m2v_free_host.restype = None
        L.m2v_register_host.argtypes = [vp, sz]
        L.m2v_unregister_host.argtypes = [vp]
        _lib = L
    return _lib


def clamp16(size16, L):
    """RTL:985-991."""
    return (1 << L) if size16 > (1 << L) else 4 if size16 < 4 else size16


def sequence_header(mbw, mbh):
    b = (C.c_uint8 * 34)()
    rc = lib().m2v_sequence_header(mbw, mbh, b)
    if rc:
        raise M2VError(rc)
    return bytes(b)


def finish_stream(data):
    """header+bodies -> complete stream: end code + zero padding (RTL:2621-2628, 2932-2937)."""
    n = len(data)
    cap = 32 * ((n + 4) // 32 + 1)
    buf = (C.c_uint8 * cap).from_buffer_copy(data + bytes(cap - n))
    tot = C.c_size_t(0)
    rc = lib().m2v_finish_stream(buf, n, cap, C.byref(tot))
    if rc:
        raise M2VError(rc)
    return bytes(buf[:tot.value])


class PinnedArray:
    """uint8 numpy view over pinned host memory from m2v_alloc_host (frames to push, or a sink for m2v_drain)."""

    def __init__(self, nbytes):
        self.ptr = lib().m2v_alloc_host(nbytes)
        if not self.ptr:
            raise M2VError(M2V_ENOMEM, 'm2v_alloc_host(%d)' % nbytes)
        self.array = np.ctypeslib.as_array((C.c_uint8 * nbytes).from_address(self.ptr))

    def close(self):
        if getattr(self, 'ptr', None):
            self.array = None
            lib().m2v_free_host(self.ptr)
            self.ptr = None

    __del__ = close


class Mpeg2Encoder:
    """One instance of the reference module: parameters are fixed at construction (RTL:11-14).  ndev > 1 spreads the
    instance over devices 0..ndev-1 of this process (m2v_create_multi); the stream is the same."""

    def __init__(self, XL=6, YL=6, VECTOR_LEVEL=3, Q_LEVEL=2, ndev=1, force_multi=False):
        self._h = C.c_void_p()
        self.XL, self.YL, self.VECTOR_LEVEL, self.Q_LEVEL = XL, YL, VECTOR_LEVEL, Q_LEVEL
        if ndev == 1 and not force_multi:
            rc = lib().m2v_create(XL, YL, VECTOR_LEVEL, Q_LEVEL, C.byref(self._h))
        else:
            rc = lib().m2v_create_multi(ndev, XL, YL, VECTOR_LEVEL, Q_LEVEL, C.byref(self._h))
        if rc:
            self._h = None
            raise M2VError(rc, 'm2v_create (a B200 / sm_100 device is required; there is no CPU fallback)')
        self.mbw = self.mbh = 0
        self.pframes_count = 0

    def close(self):
        if getattr(self, '_h', None):
            lib().m2v_destroy(self._h)
            self._h = None

    __del__ = close

    def _ck(self, rc):
        if rc < 0:
            raise M2VError(rc, (lib().m2v_last_error(self._h) or b'').decode())
        return rc

    # ---- streaming contract ----
    def begin(self, i_xsize16, i_ysize16, i_pframes_count):
        a, b = C.c_int(0), C.c_int(0)
        self._ck(lib().m2v_begin(self._h, i_xsize16, i_ysize16, i_pframes_count, C.byref(a), C.byref(b)))
        self.mbw, self.mbh, self.pframes_count = a.value, b.value, i_pframes_count
        return self.mbw, self.mbh

    def push4(self, Y, U, V):
        y = (C.c_uint8 * 4)(*Y); u = (C.c_uint8 * 4)(*U); v = (C.c_uint8 * 4)(*V)
        self._ck(lib().m2v_push4(self._h, y, u, v))

    def push_frames(self, frames):
        f = np.ascontiguousarray(frames, dtype=np.uint8)
        assert f.ndim == 4 and f.shape[1:] == (3, self.mbh * 16, self.mbw * 16), f.shape
        self._ck(lib().m2v_push_frames(self._h, f.ctypes.data, f.shape[0]))

    def sequence_stop(self):
        self._ck(lib().m2v_stop(self._h))

    @property
    def sequence_busy(self):
        return bool(lib().m2v_busy(self._h))

    def pull(self):
        """-> (32 bytes, last) or None"""
        w = (C.c_uint8 * 32)(); last = C.c_int(0)
        rc = self._ck(lib().m2v_pull(self._h, w, C.byref(last)))
        return (bytes(w), bool(last.value)) if rc == 1 else None

    def drain_into(self, buf):
        """m2v_drain straight into a caller-owned uint8 numpy buffer (no intermediate copies): -> (bytes written, last)"""
        n = C.c_size_t(0); l = C.c_int(0)
        self._ck(lib().m2v_drain(self._h, buf.ctypes.data, buf.size, C.byref(n), C.byref(l)))
        return n.value, bool(l.value)

    def drain(self, cap=1 << 24):
        """all available whole words -> (bytes, last)"""
        chunks = []
        last = False
        if getattr(self, '_drain_buf', None) is None or self._drain_buf.size < cap:
            self._drain_buf = np.empty(cap, np.uint8)
        buf = self._drain_buf
        while True:
            n = C.c_size_t(0); l = C.c_int(0)
            self._ck(lib().m2v_drain(self._h, buf.ctypes.data, cap, C.byref(n), C.byref(l)))
            chunks.append(buf[:n.value].tobytes())
            last = last or bool(l.value)
            if n.value < cap // 32 * 32:
                break
        return b''.join(chunks), last

    def encode_sequence(self, frames, i_pframes_count, partial_px4=0):
        """Replay of the testbench stimulus (TB:206-266) for one sequence: frames [n,3,H,W] uint8
        yuv444p; if partial_px4 > 0 only that many 4-pixel groups of the LAST frame are pushed
        before i_sequence_stop.  Returns the complete stream (all o_en words)."""
        f = np.ascontiguousarray(frames, dtype=np.uint8)
        n, _, H, W = f.shape
        self.begin(W // 16, H // 16, i_pframes_count)
        if partial_px4 > 0:
            if n > 1:
                self.push_frames(f[:n - 1])
            last = f[n - 1]
            yy, uu, vv = last[0].reshape(-1), last[1].reshape(-1), last[2].reshape(-1)
            for i in range(partial_px4):
                self.push4(yy[4 * i:4 * i + 4], uu[4 * i:4 * i + 4], vv[4 * i:4 * i + 4])
        else:
            self.push_frames(f)
        self.sequence_stop()
        data, last = self.drain()
        assert last and not self.sequence_busy
        return data

    # ---- device-resident bulk path ----
    def encode_gops_device(self, dev_ptr, nframes, n0, mbw, mbh, pframes_count):
        """frames already in device memory -> (device pointer of the body, length)"""
        d = C.c_void_p(); n = C.c_size_t(0)
        self._ck(lib().m2v_encode_gops_device(self._h, mbw, mbh, pframes_count, C.c_void_p(dev_ptr), nframes, n0,
                                              C.byref(d), C.byref(n)))
        return d.value, n.value

    def encode_gops_host(self, dev_ptr, nframes, n0, mbw, mbh, pframes_count, out):
        """same, body copied into the numpy uint8 array `out`; returns its length"""
        n = C.c_size_t(0)
        self._ck(lib().m2v_encode_gops_host(self._h, mbw, mbh, pframes_count, C.c_void_p(dev_ptr), nframes, n0,
                                            out.ctypes.data, out.nbytes, C.byref(n)))
        return n.value

    # ---- asynchronous chunks (two slots): submit -> size (known after the scans) -> fetch (device->host) -> wait ----
    def gops_submit(self, dev_ptr, nframes, n0, mbw, mbh, pframes_count, slot):
        self._ck(lib().m2v_gops_submit(self._h, mbw, mbh, pframes_count, C.c_void_p(dev_ptr), nframes, n0, slot))

    def gops_size(self, slot):
        n = C.c_size_t(0)
        self._ck(lib().m2v_gops_size(self._h, slot, C.byref(n)))
        return n.value

    def gops_fetch(self, slot, host_ptr, nbytes):
        self._ck(lib().m2v_gops_fetch(self._h, slot, C.c_void_p(host_ptr), nbytes))

    def gops_wait(self, slot):
        self._ck(lib().m2v_gops_wait(self._h, slot))

    def gops_body(self, slot):
        d = C.c_void_p()
        self._ck(lib().m2v_gops_body(self._h, slot, C.byref(d)))
        return d.value

    @property
    def device_count(self):
        return lib().m2v_device_count(self._h)

    def set_body_reserve(self, bytes_per_macroblock):
        """test knob (m2v_set_body_reserve): body bytes reserved per macroblock; a tiny value forces the regrow-and-rerun path"""
        self._ck(lib().m2v_set_body_reserve(self._h, int(bytes_per_macroblock)))

    def debug_copy(self, count):
        info = np.zeros(count, np.uint32); coefs = np.zeros((count, 6, 64), np.int16)
        self._ck(lib().m2v_debug_copy(self._h, info.ctypes.data, coefs.ctypes.data, count))
        return info, coefs

    @property
    def launch_count(self):
        return lib().m2v_launch_count(self._h)

    def set_timing(self, on):
        self._ck(lib().m2v_set_timing(self._h, int(on)))

    def set_limits(self, batch_frames=0, chunk_frames=0):
        """test knob (m2v_set_limits): frames per streaming batch / per encode_gops chunk; 0 = automatic"""
        self._ck(lib().m2v_set_limits(self._h, int(batch_frames), int(chunk_frames)))

    def kernel_ms(self):
        a = (C.c_float * 5)()
        self._ck(lib().m2v_kernel_ms(self._h, a))
        return list(a)

"""A from-the-standard MPEG-2 video DECODER for the syntax subset the encoder emits (ISO/IEC 13818-2 clause 7:
inverse scan 7.3, inverse quantisation 7.4 incl. saturation and mismatch control, IDCT as the ideal real-valued
transform of Annex A rounded to nearest, frame motion compensation 7.6 with the standard's half-sample rounding
and its `/ 2` truncating chroma vectors, 4:2:0).  TEST INFRASTRUCTURE - the conformance / quality harness of
SURVEY.md 8(f)-3: it reconstructs pictures from a stream with none of the encoder's own arithmetic, so that

  * the stream can be shown to decode to (nearly) what the encoder believes it reconstructed - the encoder's
    feedback loop uses the RTL's non-standard roundings (mean4 +1, floor chroma vectors,
    no mismatch control, Chen-Wang IDCT), so a conformant decoder drifts a little inside a GOP; and
  * PSNR against the source can be compared with the figure the reference publishes (README.md:748).

Built on tests/mpeg2_parser.py (which recovers macroblock types, vectors and levels from the bits).
"""
import numpy as np

import mpeg2_parser as mp
import gen_tables as G

_ZZ = np.asarray(G.ZIGZAG, dtype=np.int64).reshape(64)          # ZIGZAG[i*8+j] = scan position of coefficient (i, j)
_WI = np.asarray(G.INTRA_Q, dtype=np.int64).reshape(8, 8)       # default intra matrix (7.4.2; the stream loads none)
_C = np.array([[(np.sqrt(0.125) if u == 0 else 0.5) * np.cos((2 * x + 1) * u * np.pi / 16) for x in range(8)] for u in range(8)])


def _idct(F):
    """Annex A: ideal 8x8 IDCT on [..., v, u], rounded to nearest (halves away from zero), saturated to [-256, 255]."""
    f = np.einsum('vy,...vu,ux->...yx', _C, F.astype(np.float64), _C)
    r = np.sign(f) * np.floor(np.abs(f) + 0.5)
    return np.clip(r, -256, 255).astype(np.int64)


def _dequant(levels, intra, qscale):
    """levels [nmb, 6, 64] in scan order -> coefficients [nmb, 6, 8, 8] (7.3, 7.4)."""
    QF = levels[:, :, _ZZ].reshape(levels.shape[0], 6, 8, 8).astype(np.int64)      # QF[v][u] = scan[zigzag[v][u]]
    k = np.where(intra[:, None, None, None], 0, np.sign(QF))
    W = np.where(intra[:, None, None, None], _WI[None, None], 16)
    num = (2 * QF + k) * W * qscale
    F = np.sign(num) * (np.abs(num) // 32)                                         # '/' truncates toward zero
    dc = 2 * QF[:, :, 0, 0]                                                        # intra_dc_precision 10 bits: intra_dc_mult = 2
    F[:, :, 0, 0] = np.where(intra[:, None], dc, F[:, :, 0, 0])
    F = np.clip(F, -2048, 2047)
    s = F.sum(axis=(2, 3))                                                         # 7.4.4 mismatch control
    even = (s & 1) == 0
    l = F[:, :, 7, 7]
    F[:, :, 7, 7] = np.where(even, np.where(l & 1, l - 1, l + 1), l)
    return F


def _mc(ref, y0, x0, n, mvy, mvx):
    """7.6.4 prediction of an n x n block at (y0, x0) with a half-sample vector."""
    iy, ix, hy, hx = mvy >> 1, mvx >> 1, mvy & 1, mvx & 1
    a = ref[y0 + iy:y0 + iy + n + 1, x0 + ix:x0 + ix + n + 1].astype(np.int64)
    if a.shape != (n + 1, n + 1):                                                 # at the bottom / right border the +1 row/column is only
        a = np.pad(a, ((0, n + 1 - a.shape[0]), (0, n + 1 - a.shape[1])), mode='edge')   # touched when the half flag is set (never there)
    if hy and hx:
        return (a[:n, :n] + a[:n, 1:] + a[1:, :n] + a[1:, 1:] + 2) >> 2
    if hx:
        return (a[:n, :n] + a[:n, 1:] + 1) >> 1
    if hy:
        return (a[:n, :n] + a[1:, :n] + 1) >> 1
    return a[:n, :n]


def decode(data):
    """stream bytes -> dict(width, height, frames=[(Y, U, V) uint8 planes, 4:2:0] in display (= coding) order,
    plus the parser's picture list under 'pictures')."""
    s = mp.parse(data)
    W, H = s['width'], s['height']
    mbw, mbh = W // 16, H // 16
    frames = []
    ref = None
    for pic in s['pictures']:
        nmb = mbw * mbh
        mbs = pic['mbs']
        assert len(mbs) == nmb
        intra = np.array([m['type'] == 'intra' for m in mbs])
        lv = np.array([m['levels'] for m in mbs], dtype=np.int64)
        qscale = 2 * mbs[0]['qsc']                                                 # q_scale_type = 0
        assert all(m['qsc'] == mbs[0]['qsc'] for m in mbs)
        res = _idct(_dequant(lv, intra, qscale))                                   # [nmb, 6, 8, 8]
        Y = np.zeros((H, W), np.int64); U = np.zeros((H // 2, W // 2), np.int64); V = np.zeros((H // 2, W // 2), np.int64)
        for i, m in enumerate(mbs):
            by, bx = divmod(i, mbw)
            y0, x0 = 16 * by, 16 * bx
            if intra[i]:
                py = np.zeros((16, 16), np.int64); pu = pv = np.zeros((8, 8), np.int64)
            else:
                assert pic['type'] == 2 and ref is not None
                mvx, mvy = m['mv']
                py = _mc(ref[0], y0, x0, 16, mvy, mvx)
                cx, cy = int(mvx / 2), int(mvy / 2)                                # 7.6.3.7: '/' truncates toward zero
                pu = _mc(ref[1], y0 // 2, x0 // 2, 8, cy, cx)
                pv = _mc(ref[2], y0 // 2, x0 // 2, 8, cy, cx)
            r = res[i]
            Y[y0:y0 + 8, x0:x0 + 8] = py[:8, :8] + r[0]; Y[y0:y0 + 8, x0 + 8:x0 + 16] = py[:8, 8:] + r[1]
            Y[y0 + 8:y0 + 16, x0:x0 + 8] = py[8:, :8] + r[2]; Y[y0 + 8:y0 + 16, x0 + 8:x0 + 16] = py[8:, 8:] + r[3]
            U[y0 // 2:y0 // 2 + 8, x0 // 2:x0 // 2 + 8] = pu + r[4]
            V[y0 // 2:y0 // 2 + 8, x0 // 2:x0 // 2 + 8] = pv + r[5]
        ref = tuple(np.clip(p, 0, 255).astype(np.uint8) for p in (Y, U, V))
        frames.append(ref)
    return {'width': W, 'height': H, 'frames': frames, 'pictures': s['pictures']}


def psnr(a, b):
    mse = np.mean((np.asarray(a, np.float64) - np.asarray(b, np.float64)) ** 2)
    return 100.0 if mse == 0 else min(100.0, 10 * np.log10(255.0 ** 2 / mse))     # identical pictures: capped at 100 dB

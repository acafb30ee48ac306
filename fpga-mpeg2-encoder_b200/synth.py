"""Seeded synthetic yuv444p clips: S1 "pan", S2 "white", S3 "dark/flat",
S4 "edges".  numpy for the parity tests; a torch variant of S1 builds large clips directly in HBM
for bench.py.  All return uint8 arrays [n, 3, H, W] (Y, U, V planes per frame, as TB:210-218 reads
them)."""
import numpy as np


def _blur_field(rng, h, w, sigma, lo, hi):
    from scipy.ndimage import gaussian_filter
    f = gaussian_filter(rng.random((h, w)), sigma, mode='wrap')
    f = (f - f.min()) / max(f.max() - f.min(), 1e-9)
    return lo + f * (hi - lo)


def s1_canvas(seed, W, H, margin=32):
    rng = np.random.default_rng(seed)
    ch, cw = H + margin, W + margin
    return np.stack([_blur_field(rng, ch, cw, 3, 16, 235), _blur_field(rng, ch, cw, 6, 16, 240),
                     _blur_field(rng, ch, cw, 6, 16, 240)]).astype(np.float32)


def s1_offsets(seed, n, margin=32, step=5):
    rng = np.random.default_rng(seed + 1)
    off = np.zeros((n, 2), np.int64)
    pos = np.array([margin // 2, margin // 2])
    for t in range(n):
        off[t] = pos
        pos = np.clip(pos + rng.integers(-step, step + 1, 2), 0, margin)
    return off


def s1_pan(seed, n, W, H, step=5, noise=2):
    """Blurred-noise scene panning by a random walk of <= `step` px/frame plus +-`noise` iid noise."""
    canvas = s1_canvas(seed, W, H)
    off = s1_offsets(seed, n, step=step)
    rng = np.random.default_rng(seed + 2)
    out = np.empty((n, 3, H, W), np.uint8)
    for t in range(n):
        oy, ox = off[t]
        fr = canvas[:, oy:oy + H, ox:ox + W] + rng.integers(-noise, noise + 1, (3, H, W))
        out[t] = np.clip(np.rint(fr), 0, 255).astype(np.uint8)
    return out


def s2_white(seed, n, W, H):
    """iid uniform bytes: SAD >= 4096 everywhere -> intra macroblocks in P-frames, escape codes."""
    return np.random.default_rng(seed).integers(0, 256, (n, 3, H, W), dtype=np.uint8)


def s3_dark(seed, n, W, H):
    """Y in [0,15] blocks and constant frames: exercises the intra-cost quirk (sum < 4096),
    'MC not coded' macroblocks and DC-only tiles."""
    rng = np.random.default_rng(seed)
    out = np.empty((n, 3, H, W), np.uint8)
    for t in range(n):
        if t % 3 == 2:
            out[t, 0] = rng.integers(0, 256); out[t, 1] = rng.integers(0, 256); out[t, 2] = rng.integers(0, 256)
        else:
            blk = rng.integers(0, 16, (H // 16, W // 16), dtype=np.uint8)
            out[t, 0] = np.kron(blk, np.ones((16, 16), np.uint8)) + rng.integers(0, 2, (H, W), dtype=np.uint8) * (t % 2)
            out[t, 1] = 128 + rng.integers(-1, 2, (H, W))
            out[t, 2] = 128
    return out


def s4_edges(seed, n, W, H):
    """S1 with the motion pinned at +-6 px/frame (and sub-pixel via 2x-blended shifts) to hit the
    mv == +-YR half-pel guard and the frame-border rule."""
    canvas = s1_canvas(seed, W, H, margin=64)
    out = np.empty((n, 3, H, W), np.uint8)
    pos = np.array([32, 32]); d = np.array([6, -6])
    for t in range(n):
        oy, ox = pos
        a = canvas[:, oy:oy + H, ox:ox + W]
        if t % 2:                                   # half-pel displaced frame
            a = 0.5 * (a + canvas[:, oy:oy + H, ox + 1:ox + 1 + W])
        out[t] = np.clip(np.rint(a), 0, 255).astype(np.uint8)
        nxt = pos + d
        for k in range(2):
            if nxt[k] < 0 or nxt[k] > 64 - 1:
                d[k] = -d[k]
        pos = np.clip(pos + d, 0, 63)
    return out


GENERATORS = {'S1': s1_pan, 'S2': s2_white, 'S3': s3_dark, 'S4': s4_edges}


def s1_pan_torch(seed, n, W, H, device, out=None, chunk=16):
    """S1 built directly in device memory (uint8 [n,3,H,W]); same canvas and motion as s1_pan, noise
    from torch's generator (so NOT byte-identical to the numpy variant - parity checks copy frames
    back to the host)."""
    import torch
    canvas = torch.from_numpy(s1_canvas(seed, W, H)).to(device)
    off = s1_offsets(seed, n)
    gen = torch.Generator(device=device); gen.manual_seed(seed + 2)
    if out is None:
        out = torch.empty((n, 3, H, W), dtype=torch.uint8, device=device)
    for t in range(n):
        oy, ox = int(off[t, 0]), int(off[t, 1])
        fr = canvas[:, oy:oy + H, ox:ox + W] + torch.randint(-2, 3, (3, H, W), generator=gen, device=device)
        out[t] = fr.round_().clamp_(0, 255).to(torch.uint8)
    return out

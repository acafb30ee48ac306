"""Closed-GOP sharding across ranks (one process per GPU) and the gather of per-rank byte streams.

Why this is legal: every GOP starts with an I-frame that ignores the reference
frame (RTL:1820-1825), its header is byte aligned and says closed_gop=1 (RTL:2645-2656), slice
predictors reset per slice (RTL:2713-2715), and the only cross-GOP state - the time code - is a
closed form of the absolute frame index (RTL:2685-2698).  So rank r encodes a contiguous block of
whole GOPs with no data-path collective; NCCL (or gloo in the CPU tests) is used only to gather the
finished byte streams on rank 0, which concatenates
    [34-byte sequence header][body rank 0][body rank 1]...[00 00 01 B7][zero pad]      (RTL:2596-2628, 2932-2937).
"""


def gop_partition(nframes, pframes_count, world):
    """Contiguous, balanced blocks of whole GOPs: returns [(n0, n1)] per rank (may be empty)."""
    gop = pframes_count + 1
    ngops = (nframes + gop - 1) // gop
    base, extra = divmod(ngops, world)
    out, g = [], 0
    for r in range(world):
        k = base + (1 if r < extra else 0)
        out.append((min(g * gop, nframes), min((g + k) * gop, nframes)))
        g += k
    return out


def gather_bodies(body, dist=None, device=None):
    """body: 1-D uint8 torch tensor (this rank's bytes, on `device`).  Returns on rank 0 the list of
    per-rank uint8 tensors in rank order (None elsewhere).  Two collectives: all_gather of the byte
    counts, then gather of the payloads padded to the maximum count."""
    import torch
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return [body]
    world, rank = dist.get_world_size(), dist.get_rank()
    device = body.device if device is None else device
    n = torch.tensor([body.numel()], dtype=torch.int64, device=device)
    sizes = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(sizes, n)
    sizes = [int(s.item()) for s in sizes]
    cap = max(max(sizes), 1)
    pad = torch.zeros(cap, dtype=torch.uint8, device=device)
    pad[:body.numel()] = body
    # all_gather of the padded payloads (gather is not implemented by every NCCL build); the
    # payload is ~0.07 B/pixel, so the redundancy is irrelevant next to the encode itself
    bufs = [torch.empty(cap, dtype=torch.uint8, device=device) for _ in range(world)]
    dist.all_gather(bufs, pad)
    if rank != 0:
        return None
    return [bufs[r][:sizes[r]] for r in range(world)]


def assemble_stream(seq_header, bodies, finish):
    """rank 0: header + bodies + tail.  `finish` = package.finish_stream."""
    data = bytes(seq_header) + b''.join(bytes(b.cpu().numpy().tobytes()) if hasattr(b, 'cpu') else bytes(b) for b in bodies)
    return finish(data)

"""Closed-GOP sharding across ranks (one process per GPU) and the ways the per-rank byte streams come together.

Why this is legal: every GOP starts with an I-frame that ignores the reference
frame (RTL:1820-1825), its header is byte aligned and says closed_gop=1 (RTL:2645-2656), slice
predictors reset per slice (RTL:2713-2715), and the only cross-GOP state - the time code - is a
closed form of the absolute frame index (RTL:2685-2698).  So a rank encodes whole GOPs with no data-path
collective, and the stream is
    [34-byte sequence header][GOP 0][GOP 1]...[00 00 01 B7][zero pad]      (RTL:2596-2628, 2932-2937).

Two ways to assemble it on rank 0:
  * HostArena (the fast one, used by bench.py's `value_to_host`): a POSIX shared-memory arena mapped and pinned by
    every rank of the node; a rank learns the sizes of the others through a table in the arena and copies its body
    device->host STRAIGHT to its final offset of the concatenated stream - N PCIe links in parallel, no collective,
    no host-side concatenation copy.
  * gather_bodies (NCCL / gloo): byte counts in one all-gather, payloads with one grouped send/recv to rank 0
    (ncclSend/ncclRecv inside one group); rank 0 then holds the bodies in ITS device memory.
"""
import mmap
import os
import time

import numpy as np


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


def chunk_schedule(frames_per_rank, pframes_count, world, chunks, tail_gops=0):
    """Block-cyclic deal of one long sequence for the pipelined gather: the sequence is cut into `chunks * world` blocks of
    whole GOPs and block c*world + r goes to rank r as its chunk c.  All bodies of chunk row c are known (and on their way
    to the host) while the ranks encode chunk row c+1, so only the LAST row's device->host copy is exposed: with
    tail_gops > 0 (and chunks > 1) that last chunk is made small (tail_gops GOPs) and the others share the rest.
    Returns per rank the list of (first local frame, frames, absolute index of the first frame)."""
    gop = pframes_count + 1
    gops = (frames_per_rank + gop - 1) // gop
    tail_frames = gops * gop - frames_per_rank                   # a trailing partial GOP (end of the sequence): single rank only
    assert tail_frames == 0 or world == 1, 'only the last rank of a sequence can hold a partial GOP'
    chunks = max(1, min(chunks, gops))
    if tail_gops > 0 and chunks > 1 and gops > tail_gops + (chunks - 2):
        base, extra = divmod(gops - tail_gops, chunks - 1)
        per = [(base + (1 if c < extra else 0)) * gop for c in range(chunks - 1)] + [tail_gops * gop]
    else:
        base, extra = divmod(gops, chunks)
        per = [(base + (1 if c < extra else 0)) * gop for c in range(chunks)]
    per[-1] -= tail_frames
    out = [[] for _ in range(world)]
    n_abs = 0
    for c in range(chunks):
        for r in range(world):
            out[r].append((sum(per[:c]), per[c], n_abs))
            n_abs += per[c]
    return out


def gather_bodies(body, dist=None, device=None):
    """body: 1-D uint8 torch tensor (this rank's bytes, on `device`).  Returns on rank 0 the list of per-rank uint8
    tensors in rank order (None elsewhere).  One all-gather of the byte counts (read back with ONE device->host copy),
    then one group of point-to-point operations: every rank sends its exact payload to rank 0."""
    import torch
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return [body]
    world, rank = dist.get_world_size(), dist.get_rank()
    device = body.device if device is None else device
    n = torch.tensor([body.numel()], dtype=torch.int64, device=device)
    sizes = torch.zeros(world, dtype=torch.int64, device=device)
    try:
        dist.all_gather_into_tensor(sizes, n)
    except Exception:                                            # backends without the tensor form
        parts = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
        dist.all_gather(parts, n)
        sizes = torch.cat(parts)
    sizes = sizes.tolist()
    if rank == 0:
        bufs = [body] + [torch.empty(sizes[r], dtype=torch.uint8, device=device) for r in range(1, world)]
        ops = [dist.P2POp(dist.irecv, bufs[r], r) for r in range(1, world) if sizes[r]]
    else:
        bufs = None
        ops = [dist.P2POp(dist.isend, body, 0)] if body.numel() else []
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return bufs


def assemble_stream(seq_header, bodies, finish):
    """rank 0: header + bodies + tail.  `finish` = package.finish_stream."""
    data = bytes(seq_header) + b''.join(bytes(b.cpu().numpy().tobytes()) if hasattr(b, 'cpu') else bytes(b) for b in bodies)
    return finish(data)


class HostArena:
    """Shared, pinned host memory of one node: [size table | concatenated stream].  Every rank maps /dev/shm/<name> and pins
    it through the library (m2v_register_host), so that m2v_gops_fetch can copy a body device->host to any offset of it.

    Size table: entry (row, rank) = (epoch, bytes) as two int64; a rank publishes `bytes` then `epoch` (x86 stores are
    ordered; the reader spins on the epoch), so no collective and no system call sits between "my scan is done" and "I
    know where my body goes"."""
    TABLE = 1 << 16

    def __init__(self, pkg, name, nbytes, rank, world, rows=64, pin=True, directory='/dev/shm'):
        self.pkg, self.rank, self.world, self.rows, self.pinned = pkg, rank, world, rows, pin
        self.nbytes = self.TABLE + nbytes
        assert rows * world * 16 + world * 16 <= self.TABLE
        self.mm = self._pin = None
        if directory is None:                                     # one process: private pinned memory, nothing to share
            assert world == 1
            self._pin = pkg.PinnedArray(self.nbytes)
            self.buf = self._pin.array
            self.buf[:self.TABLE] = 0
            self.addr = self._pin.ptr
            self.pinned = False
            self._views()
            return
        self.path = os.path.join(directory, name)
        if rank == 0:                                             # callers put a barrier between rank 0's constructor and the others'
            fd = os.open(self.path, os.O_CREAT | os.O_RDWR | os.O_TRUNC, 0o600)
            os.ftruncate(fd, self.nbytes)
        else:
            fd = os.open(self.path, os.O_RDWR)
            assert os.fstat(fd).st_size == self.nbytes
        self.mm = mmap.mmap(fd, self.nbytes)
        os.close(fd)
        self.buf = np.frombuffer(self.mm, dtype=np.uint8)
        self.addr = self.buf.ctypes.data
        if pin:                                                   # pin=False: host-only use (the CPU tests of the table and the barrier)
            rc = pkg.lib().m2v_register_host(self.addr, self.nbytes)
            if rc:
                raise pkg.M2VError(rc, 'm2v_register_host(host arena)')
        self._views()

    def _views(self):
        rows, world = self.rows, self.world
        self.table = self.buf[:rows * world * 16].view(np.int64).reshape(rows, world, 2)
        self.flags = self.buf[rows * world * 16:rows * world * 16 + world * 16].view(np.int64).reshape(world, 2)
        self.stream = self.buf[self.TABLE:]
        self.stream_addr = self.addr + self.TABLE

    def publish(self, row, epoch, nbytes):
        self.table[row % self.rows, self.rank, 1] = nbytes
        self.table[row % self.rows, self.rank, 0] = epoch

    def sizes(self, row, epoch, timeout=30.0):
        """byte counts of every rank for this row (spins until all have published this epoch)"""
        t = self.table[row % self.rows]
        t0 = time.perf_counter()
        while not bool((t[:, 0] == epoch).all()):
            if time.perf_counter() - t0 > timeout:
                raise TimeoutError('host arena: a rank did not publish row %d epoch %d' % (row, epoch))
        return t[:, 1].copy()

    def barrier(self, epoch, timeout=120.0, sleep=0.0):
        """all ranks have reached `epoch` (monotonically increasing); sleep > 0 polls instead of spinning (long waits)"""
        self.flags[self.rank, 0] = epoch
        t0 = time.perf_counter()
        while not bool((self.flags[:, 0] >= epoch).all()):
            if sleep:
                time.sleep(sleep)
            if time.perf_counter() - t0 > timeout:
                raise TimeoutError('host arena: barrier %d' % epoch)

    def close(self):
        if getattr(self, '_pin', None) is not None:
            self.table = self.flags = self.stream = self.buf = None
            self._pin.close(); self._pin = None
            return
        if getattr(self, 'mm', None) is None:
            return
        if self.pinned:
            self.pkg.lib().m2v_unregister_host(self.addr)
        self.table = self.flags = self.stream = self.buf = None
        try:
            self.mm.close()
        except BufferError:
            pass
        self.mm = None
        if self.rank == 0:
            try:
                os.unlink(self.path)
            except OSError:
                pass

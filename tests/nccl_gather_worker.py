"""Worker of tests/test_host_engine_gpu.py::test_nccl_gather_two_ranks (run under torch.distributed.run, one rank per GPU).

1. NCCL: contiguous GOP blocks per rank, sharding.gather_bodies (all-gather of sizes + grouped send/recv to rank 0), rank 0
   assembles header + bodies + tail and compares with the oracle.
2. Host arena: block-cyclic chunks (sharding.chunk_schedule), every rank copies its bodies device->host straight to their final
   offsets of the shared pinned arena (m2v_gops_submit / size / fetch), rank 0 compares the arena with the oracle.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def main():
    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    import oracle_binding as ob
    pkg = ge.load_package(); synth = ge.load_synth()
    from fpga_mpeg2_encoder_b200 import sharding
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    W, H, P, n = 160, 96, 3, 50
    mbw, mbh = W // 16, H // 16
    fr = synth.s1_pan(71, n, W, H)                                  # every rank builds the same clip and keeps its share on its GPU
    want = ob.encode(fr, mbw, mbh, P, XL=6, YL=6) if rank == 0 else None
    enc = pkg.Mpeg2Encoder(XL=6, YL=6)
    fsz = 3 * W * H

    # ---- 1. NCCL gather ----
    n0, n1 = sharding.gop_partition(n, P, world)[rank]
    d = torch.from_numpy(fr[n0:n1]).to(dev)
    torch.cuda.synchronize()
    ptr, ln = enc.encode_gops_device(d.data_ptr(), n1 - n0, n0, mbw, mbh, P)

    class _DevView:
        def __init__(self, p, k):
            self.__cuda_array_interface__ = {'shape': (k,), 'typestr': '|u1', 'data': (p, False), 'version': 3}
    body = torch.as_tensor(_DevView(ptr, ln), device=dev).clone()
    bodies = sharding.gather_bodies(body, dist, dev)
    if rank == 0:
        got = sharding.assemble_stream(pkg.sequence_header(mbw, mbh), bodies, pkg.finish_stream)
        assert got == want, 'NCCL gather: stream differs from the oracle (%d vs %d bytes)' % (len(got), len(want))
        print('NCCL GATHER OK %d bytes' % len(got), flush=True)
    else:
        assert bodies is None

    # ---- 2. host arena, block-cyclic chunks ----
    per_rank = 24                                                   # frames per rank: 6 GOPs in 3 chunks of 2 GOPs
    sched = sharding.chunk_schedule(per_rank, P, world, 3)
    total_frames = per_rank * world
    mine = sched[rank]
    loc = np.concatenate([fr[a:a + k] for (_, k, a) in mine])       # this rank's frames, chunk after chunk
    d2 = torch.from_numpy(loc).to(dev)
    torch.cuda.synchronize()
    name = 'm2v_test_%s' % os.environ.get('MASTER_PORT', '0')
    if rank == 0:
        arena = sharding.HostArena(pkg, name, 8 << 20, rank, world)
    dist.barrier()
    if rank != 0:
        arena = sharding.HostArena(pkg, name, 8 << 20, rank, world)
    dist.barrier()
    for step in range(2):                                           # twice: the epochs of the table must separate the steps
        if rank == 0:
            arena.stream[:34] = np.frombuffer(pkg.sequence_header(mbw, mbh), np.uint8)
        base = 34
        C = len(mine)
        f0, k0, a0 = mine[0]
        enc.gops_submit(d2.data_ptr() + f0 * fsz, k0, a0, mbw, mbh, P, 0)
        for c in range(C):
            if c + 1 < C:
                f1, k1, a1 = mine[c + 1]
                enc.gops_submit(d2.data_ptr() + f1 * fsz, k1, a1, mbw, mbh, P, (c + 1) & 1)
            nb = enc.gops_size(c & 1)
            epoch = step * C + c + 1
            arena.publish(c, epoch, nb)
            sz = arena.sizes(c, epoch)
            off = base + int(sz[:rank].sum())
            enc.gops_fetch(c & 1, arena.stream_addr + off, nb)
            base += int(sz.sum())
        enc.gops_wait(0); enc.gops_wait(1)
        arena.barrier(step + 1)
        if rank == 0:
            got = pkg.finish_stream(arena.stream[:base].tobytes())
            ref = ob.encode(fr[:total_frames], mbw, mbh, P, XL=6, YL=6)
            assert got == ref, 'host arena: stream differs from the oracle (%d vs %d bytes)' % (len(got), len(ref))
        dist.barrier()
    if rank == 0:
        print('HOST ARENA OK', flush=True)
    arena.close(); enc.close()
    dist.barrier(); dist.destroy_process_group()


if __name__ == '__main__':
    main()

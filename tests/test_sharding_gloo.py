"""world_size-2 CPU (gloo) test of the multi-GPU plumbing: stream partition, max-over-ranks timing and
the concatenating gather.  The data path itself has no collective (streams are independent)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from wekws_b200.sharding import gather_streams, max_over_ranks, stream_slice


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, B, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b, e = stream_slice(B, rank, world)
    # each rank "scores" its own streams: posterior of stream i is a function of i only
    local = torch.arange(b, e, dtype=torch.float32).reshape(-1, 1, 1).expand(-1, 3, 2).contiguous() * 0.5
    full = gather_streams(local, B)
    tmax = max_over_ranks(1.0 + rank)
    dist.barrier()
    q.put((rank, b, e, full.numpy().tolist(), tmax))     # plain python objects: no shared-memory handles
    dist.destroy_process_group()


@pytest.mark.parametrize("B", [7, 1024])
def test_two_rank_sharding(B):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, B, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # slices tile [0, B) without gaps or overlap
    assert res[0][1] == 0 and res[0][2] == res[1][1] and res[1][2] == B
    expect = torch.arange(B, dtype=torch.float32).reshape(-1, 1, 1).expand(-1, 3, 2) * 0.5
    for r in res:
        assert torch.equal(torch.tensor(r[3]), expect)          # rank-ordered concatenation == unsharded result
        assert r[4] == 2.0                        # max over ranks


def test_slices_are_balanced():
    for B in (1, 5, 148, 1024, 10000):
        for world in (1, 2, 4, 8):
            sizes = [stream_slice(B, r, world) for r in range(world)]
            assert sizes[0][0] == 0 and sizes[-1][1] == B
            assert all(sizes[i][1] == sizes[i + 1][0] for i in range(world - 1))
            n = [e - b for b, e in sizes]
            assert max(n) - min(n) <= 1

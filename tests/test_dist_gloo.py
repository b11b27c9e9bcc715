"""CPU, world_size 2 over gloo: the host-side logic of the env-parallel build (gennbv_b200/dist.py) -- env sharding,
gradient averaging in one flat bucket, rank-consistent KL early stop, parameter broadcast."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gennbv_b200 import dist as gdist


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # shards tile the env range
        start, count = gdist.shard_envs(2048 + 3, rank, world)
        t = torch.zeros(2048 + 3)
        t[start:start + count] = 1
        dist.all_reduce(t)
        assert bool((t == 1).all())
        # flat-bucket gradient mean
        g = torch.full((1000,), float(rank + 1))
        gdist.allreduce_mean_(g)
        assert torch.allclose(g, torch.full((1000,), (1 + world) / 2 * 1.0))
        # KL consensus: only rank 1 exceeds the threshold, both ranks must stop; neither stops below it
        kl = torch.tensor([0.2 if rank == 1 else 0.01])
        assert gdist.should_stop(kl, 0.05) is True
        assert gdist.should_stop(torch.tensor([0.01]), 0.05) is False
        assert gdist.should_stop(kl, None) is False
        # lock-step loop: 5 minibatches, rank 0 would stop at 3, rank 1 never -> both run exactly 3 all-reduces
        n = 0
        for mb in range(5):
            my_kl = torch.tensor([1.0 if (rank == 0 and mb == 3) else 0.0])
            if gdist.should_stop(my_kl, 0.05):
                break
            gdist.allreduce_mean_(torch.ones(8))
            n += 1
        assert n == 3
        # broadcast of parameters + buffers
        flat = torch.full((16,), float(rank))
        bufs = [torch.full((4,), float(rank)), torch.tensor(rank, dtype=torch.int64)]
        gdist.broadcast_state_(flat, bufs)
        assert float(flat.sum()) == 0 and float(bufs[0].sum()) == 0 and int(bufs[1]) == 0
        q.put((rank, "ok"))
    except Exception as e:       # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_shard_envs_is_a_partition():
    for total, world in ((2048, 8), (257, 4), (5, 8)):
        spans = [gdist.shard_envs(total, r, world) for r in range(world)]
        assert spans[0][0] == 0 and sum(c for _, c in spans) == total
        for (s0, c0), (s1, _) in zip(spans, spans[1:]):
            assert s0 + c0 == s1
        assert max(c for _, c in spans) - min(c for _, c in spans) <= 1


def test_single_process_is_identity():
    g = torch.arange(4.0)
    assert torch.equal(gdist.allreduce_mean_(g.clone()), g) and gdist.world() == (0, 1)

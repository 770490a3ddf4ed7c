"""CPU tier: the N > 1 plumbing (tetwild_b200/shard.py) on two gloo ranks: index-range split, mesh broadcast, ragged
all-gather of per-query results, max-over-ranks timing. The compute stand-in here is the CPU oracle (the checker): the
CUDA path itself is exercised per rank by the gpu tier and by bench.py --gpus N."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tetwild_b200 import shard, synth


def test_shard_bounds_cover():
    for n in (0, 1, 7, 8, 9, 1000, 10_000_019):
        for world in (1, 2, 3, 4, 8):
            cuts = [shard.shard_bounds(n, world, r) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(cuts[r][1] == cuts[r + 1][0] for r in range(world - 1))
            sizes = [e - b for b, e in cuts]
            assert max(sizes) - min(sizes) <= 1 and sizes == shard.shard_sizes(n, world)
    with pytest.raises(ValueError):
        shard.shard_bounds(10, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle as O
        O.build()
        # rank 0 owns the mesh, everybody gets a bit-identical replica
        V0, F0 = synth.icosphere(3)
        V, F = shard.broadcast_mesh(V0 if rank == 0 else None, F0 if rank == 0 else None)
        assert V.dtype == np.float64 and F.dtype == np.uint32 and np.array_equal(V, V0) and np.array_equal(F, F0)
        sd, eps, eps2 = synth.state_eps(1e-2, diag=1.0)
        n = 10_001  # odd: the shards differ by one element
        P = synth.envelope_points(V, F, n, eps, seed=3)  # same seed on every rank = the "global" batch
        S = O.Surface(V, F)

        def local(b, e):
            return torch.from_numpy(S.points_out(P[b:e], eps2))

        got = shard.sharded_decisions(n, local)
        want = S.points_out(P, eps2)
        assert got.dtype == torch.uint8 and np.array_equal(got.numpy(), want)
        # 8-byte values, 2-D payload, empty shards
        b, e = shard.shard_bounds(n, world, rank)
        W = shard.all_gather_ragged(torch.from_numpy(P[b:e]), n)
        assert np.array_equal(W.numpy(), P)
        assert shard.all_gather_ragged(torch.empty((0, 3)), 0).shape == (0, 3)
        one = shard.all_gather_ragged(torch.full((shard.shard_sizes(1, world)[rank],), 7, dtype=torch.uint8), 1)  # one rank holds nothing
        assert one.tolist() == [7]
        assert shard.max_over_ranks(1.0 + rank) == float(world)
        q.put((rank, "ok"))
    except Exception as ex:  # noqa: BLE001
        q.put((rank, "FAIL %r" % (ex,)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_ranks_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(30)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res

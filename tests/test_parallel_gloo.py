"""Host-side logic of the multi-GPU path on CPU: utterance sharding and the score all-gather with the gloo backend,
world_size 2 and 3 (ragged)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from avzoom import parallel
    lo, hi = parallel.shard_range(n_total, rank, world)
    # a stand-in for the per-utterance scores of this shard: row u = (u, 2u, u^2, -u)
    u = torch.arange(lo, hi, dtype=torch.float32)
    local = torch.stack([u, 2 * u, u * u, -u], dim=1)
    allsc = parallel.gather_scores(local, n_total)
    ret[rank] = allsc.clone()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n_total", [(2, 16), (2, 7), (3, 10)])
def test_shards_and_score_allgather(world, n_total):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), n_total, ret), nprocs=world, join=True)
    u = torch.arange(n_total, dtype=torch.float32)
    want = torch.stack([u, 2 * u, u * u, -u], dim=1)
    for r in range(world):
        assert torch.equal(ret[r], want), f"rank {r} gathered scores out of order"


def test_shard_ranges_cover_exactly():
    from avzoom import parallel
    for n in (1, 7, 64, 1024, 65536):
        for w in (1, 2, 3, 4, 8):
            edges = [parallel.shard_range(n, r, w) for r in range(w)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(edges, edges[1:]))
            sizes = [hi - lo for lo, hi in edges]
            assert max(sizes) - min(sizes) <= 1

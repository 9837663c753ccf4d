"""N > 1 host-side path on CPU: world_size-2 gloo run of the column partition + final gather."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lightspinner_b200.sharding import gather_columns, shard_range


def test_shard_range_covers_all_columns_once():
    for ncol in (1, 2, 7, 8, 164, 165, 8192):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                s, c = shard_range(ncol, world, r)
                seen += list(range(s, s + c))
            assert seen == list(range(ncol))


def _worker(rank, world, port, ncol, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    s, c = shard_range(ncol, world, rank)
    # stand-in for the per-column emergent intensity: a deterministic function of the global column index
    local = torch.stack([torch.arange(6, dtype=torch.float64) * 0.5 + col for col in range(s, s + c)]) if c else \
        torch.empty((0, 6), dtype=torch.float64)
    out = gather_columns(local, ncol)
    q.put((rank, out.numpy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('ncol', [8, 7])
def test_gather_two_ranks_gloo(ncol):
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ncol, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = np.stack([np.arange(6) * 0.5 + col for col in range(ncol)])
    assert np.array_equal(res[0], expect) and np.array_equal(res[1], expect)

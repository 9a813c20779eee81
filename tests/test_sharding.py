"""N>1 host logic on CPU: the batched workload shards by problem index with no data-path
collective; results are only gathered.  Run with gloo, world_size 2."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import ipm_zoo_b200 as z
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = z.shard_range(total, world, rank)
    owned = torch.zeros(total, dtype=torch.int64)
    owned[lo:hi] = 1
    dist.all_reduce(owned)  # test-only check: every problem owned exactly once
    # the bench's max-over-ranks timing and sum-over-ranks work, as in bench.py
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    n = torch.tensor([float(hi - lo)], dtype=torch.float64)
    dist.all_reduce(n, op=dist.ReduceOp.SUM)
    if rank == 0:
        q.put((owned.tolist(), t.item(), n.item()))
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [4096, 7, 2])
def test_shard_by_problem_index_world2(total):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + total) % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    owned, tmax, nsum = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert owned == [1] * total
    assert tmax == 2.0 and nsum == float(total)


def test_shard_range_partition_properties():
    sys.path.insert(0, ROOT)
    import ipm_zoo_b200 as z
    for total in (0, 1, 5, 4096, 4097):
        for world in (1, 2, 4, 8):
            r = [z.shard_range(total, world, k) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == total
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1

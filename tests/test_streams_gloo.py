"""N>1 host logic on CPU: world_size-2 gloo run of the sharding + max-over-ranks timing reduce."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tracking_b200 import streams


def test_shard_streams_partition():
    for n in (1, 7, 64, 128):
        for g in (1, 2, 4, 8):
            owned = [streams.shard_streams(n, g, r) for r in range(g)]
            flat = sorted(s for o in owned for s in o)
            assert flat == list(range(n))                       # every stream exactly once
            assert max(map(len, owned)) - min(map(len, owned)) <= 1
            assert streams.streams_per_rank(n, g) == [len(o) for o in owned]
    assert streams.streams_per_rank(64, 8) == [8] * 8           # BASELINE config 4
    with pytest.raises(ValueError):
        streams.shard_streams(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = streams.shard_streams(5, world, rank)
    # what bench.py does: barrier, local timed region, max over ranks, sum of units
    dist.barrier()
    t = torch.tensor([0.010 * (rank + 1)], dtype=torch.float64)      # rank 1 is slower
    units = torch.tensor([float(len(mine) * 1000)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(units, op=dist.ReduceOp.SUM)
    if rank == 0:
        out.put((t.item(), units.item()))
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_gloo_timing_reduce():
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    t, units = q.get()
    assert abs(t - 0.020) < 1e-12 and units == 5000.0
    assert streams.aggregate_throughput([3000, 2000], [0.010, 0.020]) == 5000 / 0.020

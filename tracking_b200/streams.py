"""Camera-stream sharding across GPUs: stream s lives on GPU s mod G.

Streams are independent (a stream's model depends only on its own past frames), so ranks never
exchange data: there is no collective on the data path.  torch.distributed is used by bench.py
only for the start/stop barrier and the max-over-ranks of the timed region.
"""
from __future__ import annotations


def shard_streams(nstreams: int, world_size: int, rank: int):
    """Stream ids owned by `rank` (round robin: s mod world_size == rank)."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank/world_size")
    return list(range(rank, nstreams, world_size))


def streams_per_rank(nstreams: int, world_size: int):
    return [len(shard_streams(nstreams, world_size, r)) for r in range(world_size)]


def aggregate_throughput(units_per_rank, seconds_per_rank):
    """Whole-job throughput = all units / slowest rank's time (max over ranks)."""
    return float(sum(units_per_rank)) / max(seconds_per_rank)

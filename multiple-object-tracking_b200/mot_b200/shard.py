"""Multi-GPU plumbing for the stream-sharded path: one process per GPU, stream s -> rank s mod G,
no data-path collective (streams are independent: all tracks, frames, models and the Hungarian of a stream stay on one
GPU).  torch.distributed is used for exactly two things: the barrier around the timed region and the MAX over ranks of
the device time."""
import torch
import torch.distributed as dist


def streams_of_rank(n_streams, rank, world):
    """Global stream ids owned by `rank` (round robin, the reference's natural unit: one stream per process)."""
    return list(range(rank, n_streams, world))


def max_over_ranks(value, device=None):
    """Device time of a step is the max over ranks (never wall clock)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def total_over_ranks(value, device=None):
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())

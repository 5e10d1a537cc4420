"""Multi-GPU plumbing of the benchmark / serving harness: sessions are independent, so a job is sharded as
`sessions_per_gpu` streams per rank with NO collective on the data path. torch.distributed is used
only to line ranks up (barrier) and to combine timings (max over ranks) and counts (sum)."""
import os


def rank_info():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def shard_stream_ids(rank, world, sessions_per_rank, first_id=0):
    """Global stream ids owned by `rank` (weak scaling: every rank gets the same count)."""
    base = first_id + rank * sessions_per_rank
    return list(range(base, base + sessions_per_rank))


def combine(dist, device, max_values, sum_values):
    """max over ranks of each entry of max_values, sum over ranks of sum_values. dist=None -> identity."""
    if dist is None:
        return list(max_values), list(sum_values)
    import torch
    m = torch.tensor(list(max_values), dtype=torch.float64, device=device)
    s = torch.tensor(list(sum_values), dtype=torch.float64, device=device)
    dist.all_reduce(m, op=dist.ReduceOp.MAX)
    dist.all_reduce(s, op=dist.ReduceOp.SUM)
    return m.tolist(), s.tolist()

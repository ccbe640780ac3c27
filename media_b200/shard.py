"""media_b200/shard.py -- multi-GPU plumbing of the encode path: sessions are independent, so they are sharded across
ranks (one process per GPU) with no data-path collective; torch.distributed is used only to agree on timing and totals.
Analogue in the reference: one encoder object per session, placed per device (video_codec/VideoEncoderNetint.cpp:300-302)."""
import torch
import torch.distributed as dist


def sessions_of_rank(n_sessions, world, rank):
    """contiguous, balanced split: the first n % world ranks take one extra session"""
    base, rem = divmod(n_sessions, world)
    start = rank * base + min(rank, rem)
    return list(range(start, start + base + (1 if rank < rem else 0)))


def _dev():
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")


def max_over_ranks(v):
    if not (dist.is_available() and dist.is_initialized()):
        return float(v)
    t = torch.tensor([float(v)], dtype=torch.float64, device=_dev())
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


def sum_over_ranks(v):
    if not (dist.is_available() and dist.is_initialized()):
        return v
    t = torch.tensor([float(v)], dtype=torch.float64, device=_dev())
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return int(round(t.item()))

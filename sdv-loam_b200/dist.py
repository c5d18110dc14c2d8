"""Sequence sharding + timing reduction for the batched mode: one process per GPU, sequence i -> rank i mod N,
no data-path collective; torch.distributed (NCCL on GPUs, gloo in the CPU tests) only for barriers and the final reductions."""
from __future__ import annotations


def shard_sequences(n_sequences: int, rank: int, world: int):
    """Indices of the sequences (or Monte-Carlo seeds) rank `rank` owns."""
    return list(range(rank, n_sequences, world))


def reduce_step_report(local_frames: int, local_seconds: float, pose_digest: float, device=None):
    """All ranks -> (total frames, max seconds, sum of digests).  Max over ranks is the job time (contract: never wall-clock of one rank)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return local_frames, local_seconds, pose_digest
    t = torch.tensor([float(local_frames), float(pose_digest)], dtype=torch.float64, device=device)
    m = torch.tensor([float(local_seconds)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM); dist.all_reduce(m, op=dist.ReduceOp.MAX)
    return int(t[0].item()), float(m[0].item()), float(t[1].item())

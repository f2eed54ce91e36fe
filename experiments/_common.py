"""Shared launcher plumbing: import paths, torchrun initialisation, per-rank sharding helpers."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "continual-learning-for-dynamic-video-quality-enhancement_b200")
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)


def setup_distributed():
    """Initialise torch.distributed from torchrun's environment (no-op for a single process).
    Returns (rank, local_rank, world, device)."""
    import torch
    from nerve_cl_b200 import distributed as nd
    rank, local_rank, world = nd.init_from_env()
    if not torch.cuda.is_available():
        raise RuntimeError("the nerve_cl_b200 hot path has no CPU implementation: a CUDA device (B200) is required")
    torch.cuda.set_device(local_rank)
    return rank, local_rank, world, torch.device("cuda", local_rank)


def shard(n: int, rank: int, world: int):
    """Indices of a length-n dataset owned by `rank` (contiguous shards, remainder dropped so that every
    rank runs the same number of steps)."""
    per = n // world
    return range(rank * per, (rank + 1) * per)


def log(rank: int, *a):
    if rank == 0:
        print(*a, flush=True)

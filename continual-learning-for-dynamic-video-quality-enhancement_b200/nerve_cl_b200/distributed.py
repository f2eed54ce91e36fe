"""Data-parallel plumbing for the hot path (one process per GPU, ``torch.distributed`` over NCCL).

The reference has no distributed code at all; the path shards naturally by
sample, so training is batch-sharded with a gradient all-reduce and inference is clip-sharded with no
collective.

``GradSync`` is the bucketed gradient all-reduce, driven by the engine's own backward: the engine
reports when every gradient under a parameter-name prefix is final (``upsampler.``, ``gff.``,
``residual_blocks.7.`` ... ``feature_extractor.`` -- the reverse of registration order, so the final
region is a growing *suffix* of the flat fp32 gradient buffer) and each bucket (a contiguous slice of
that buffer) is all-reduced asynchronously as soon as it is complete, while the rest of backward keeps
the SMs busy.  c10d's NCCL backend enqueues the collective on its own stream after the work already
queued on the compute stream; ``finish()`` makes the compute stream wait for the last bucket.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

Tensor = torch.Tensor


def init_from_env(backend: Optional[str] = None) -> Tuple[int, int, int]:
    """Initialise the default process group from torchrun's environment.  Returns (rank, local_rank, world)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend, rank=rank, world_size=world,
                                    device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, local_rank, world


def plan_buckets(layout: Dict[str, Tuple[int, int, object]], bucket_bytes: int) -> List[Tuple[int, int, List[str]]]:
    """Cut the flat gradient buffer into buckets [(start, end, [param names])], last parameters first.

    ``layout`` maps name -> (offset, numel, shape) in registration order.  Buckets are contiguous element
    ranges that close once they hold >= bucket_bytes (the first bucket in backward order is kept small by
    construction: it starts with upsampler/gff, so communication starts early)."""
    items = sorted(layout.items(), key=lambda kv: kv[1][0])
    end_of = {}
    for i, (name, (off, n, _)) in enumerate(items):
        end_of[name] = items[i + 1][1][0] if i + 1 < len(items) else off + n
    buckets: List[Tuple[int, int, List[str]]] = []
    cur_names: List[str] = []
    cur_end = None
    for name, (off, n, _) in reversed(items):
        if cur_end is None:
            cur_end = end_of[name]
        cur_names.append(name)
        if (cur_end - off) * 4 >= bucket_bytes:
            buckets.append((off, cur_end, cur_names))
            cur_names, cur_end = [], None
    if cur_names:
        first_off = items[0][1][0]
        buckets.append((first_off, cur_end, cur_names))
    return buckets


class GradSync:
    """Bucketed, backward-overlapped all-reduce(mean) of the engine's flat gradient buffer."""

    def __init__(self, process_group=None, bucket_bytes: int = 1 << 20):
        self.group = process_group
        self.bucket_bytes = bucket_bytes
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self._flat: Optional[Tensor] = None
        self._buckets: List[Tuple[int, int, List[str]]] = []
        self._pending: List[set] = []
        self._works = []
        self._launched: List[bool] = []
        self.launch_log: List[Tuple[int, int]] = []     # (start, end) in launch order, for tests

    def begin(self, flat: Tensor, layout: Dict[str, Tuple[int, int, object]]) -> None:
        self._flat = flat
        self._buckets = plan_buckets(layout, self.bucket_bytes)
        self._pending = [set(names) for _, _, names in self._buckets]
        self._launched = [False] * len(self._buckets)
        self._works = []
        self.launch_log = []

    def _launch(self, i: int) -> None:
        a, b, _ = self._buckets[i]
        self._launched[i] = True
        self.launch_log.append((a, b))
        if self.world == 1:
            return
        chunk = self._flat[a:b]
        if dist.get_backend(self.group) == "nccl":
            work = dist.all_reduce(chunk, op=dist.ReduceOp.AVG, group=self.group, async_op=True)
            self._works.append((work, None))
        else:   # gloo (CPU tests, and the 2-ranks-on-one-GPU parity test): no AVG
            work = dist.all_reduce(chunk, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
            self._works.append((work, chunk))

    def on_ready(self, prefix: str) -> None:
        """All gradients whose parameter name starts with ``prefix`` are final."""
        for i, pend in enumerate(self._pending):
            if self._launched[i] or not pend:
                continue
            done = {n for n in pend if n.startswith(prefix)}
            if done:
                pend -= done
            if not pend:
                # buckets must be launched in the same order on every rank: flush predecessors first
                for j in range(i + 1):
                    if not self._launched[j] and not self._pending[j]:
                        self._launch(j)

    def finish(self) -> None:
        for i in range(len(self._buckets)):
            if not self._launched[i]:
                self._launch(i)
        for work, chunk in self._works:
            work.wait()
            if chunk is not None:
                chunk.div_(self.world)
        self._works = []
        self._flat = None


def broadcast_parameters(module: torch.nn.Module, src: int = 0, group=None) -> None:
    """Make every rank start from rank ``src``'s parameters and buffers (DDP's constructor behaviour)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src=src, group=group)


def sync_buffers(module: torch.nn.Module, src: int = 0, group=None) -> None:
    """Broadcast rank ``src``'s buffers (the BatchNorm running statistics) to every rank -- what DDP does at the start
    of every forward.  The ranks' statistics drift apart during training because BatchNorm stays per-rank (the
    single-process reference has no SyncBN); call this before validation / checkpointing so every rank evaluates,
    and rank 0 saves, the same model."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    for t in module.buffers():
        dist.broadcast(t.data, src=src, group=group)


def param_divergence(module: torch.nn.Module, group=None) -> float:
    """max over ranks and parameters of |theta_rank - theta_rank0| (must be exactly 0 under data parallelism)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return 0.0
    worst = torch.zeros(1, device=next(module.parameters()).device)
    for p in module.parameters():
        ref = p.detach().clone()
        dist.broadcast(ref, src=0, group=group)
        worst = torch.maximum(worst, (p.detach() - ref).abs().max().reshape(1))
    dist.all_reduce(worst, op=dist.ReduceOp.MAX, group=group)
    return float(worst.item())


def shard_clips(num_clips: int, rank: int, world: int) -> List[int]:
    """Clip-sharded inference (no collective): clip k runs on rank k mod world."""
    return list(range(rank, num_clips, world))


def data_parallel(module, bucket_bytes: int = 1 << 20, group=None):
    """Wire a SuperResolutionNet for data-parallel training: broadcast weights, install GradSync."""
    broadcast_parameters(module, 0, group)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        module.set_gradient_sync(GradSync(group, bucket_bytes))
    return module

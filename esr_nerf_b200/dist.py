"""Ray-sharded data parallelism for the render step: one process per GPU, each rank renders a
contiguous slice of the global ray batch against replicated parameters, gradients are summed with ONE collective
step per training step (`torch.distributed` all_reduce: NCCL over NVLink on the GPUs, gloo in the CPU tests).
No other collective is on the path.  The reference has no distributed code at all (cfg/__init__.yaml:24)."""
from __future__ import annotations

from typing import Dict, Iterable

import torch


def shard_slice(n: int, rank: int, world: int) -> slice:
    """Contiguous slice of a global batch of n rays owned by `rank` (keeps BatchSampler's contiguous-slice semantics,
    utils2/utils.py:106-119; the last global ray stays on the last rank, which matters for the reference's
    last-ray-only entropy term, fine.py:378)."""
    per = (n + world - 1) // world
    lo = min(rank * per, n)
    return slice(lo, min(lo + per, n))


def shard_batch(batch: Dict[str, torch.Tensor], rank: int, world: int) -> Dict[str, torch.Tensor]:
    n = batch["rays_o"].shape[0]
    sl = shard_slice(n, rank, world)
    return {k: (v[sl] if torch.is_tensor(v) and v.dim() > 0 and v.shape[0] == n else v) for k, v in batch.items()}


def allreduce_gradients(params: Iterable[torch.nn.Parameter], group=None) -> int:
    """Sum every parameter gradient over the ranks (losses are written as sums over rays / the global ray count, so
    a plain sum reproduces the single-process gradient).  Parameters a rank did not touch contribute zeros.
    Returns the number of bytes reduced."""
    import torch.distributed as dist

    nbytes = 0
    for p in params:
        if not p.requires_grad:
            continue
        if p.grad is None:
            p.grad = torch.zeros_like(p)
        dist.all_reduce(p.grad, op=dist.ReduceOp.SUM, group=group)
        nbytes += p.grad.numel() * p.grad.element_size()
    return nbytes

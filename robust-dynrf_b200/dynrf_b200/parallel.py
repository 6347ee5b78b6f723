"""Multi-GPU plumbing for the ray-march path: rays are independent, so the batch is
partitioned into contiguous per-rank slices with replicated parameters.  Forward needs exactly one
collective -- an all-gather of the composited per-ray maps ([N,16] floats) -- and a data-parallel
train step additionally sums the parameter gradients with one all-reduce over a flat buffer.
Backend-agnostic (NCCL on the GPUs; gloo in the CPU tests)."""
from __future__ import annotations

from typing import Iterable, List, Tuple

import torch


def shard_bounds(n_total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced slice [lo, hi) of a batch of n_total rays for `rank` (first ranks take the remainder)."""
    base, rem = divmod(n_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_ray_maps(local: torch.Tensor, world: int, counts: List[int] | None = None) -> torch.Tensor:
    """all-gather of per-ray outputs [n_local, C] -> [n_total, C] in rank order (ragged slices supported)."""
    if world == 1:
        return local
    import torch.distributed as dist

    if counts is None or len(set(counts)) == 1:
        out = torch.empty(local.shape[0] * world, *local.shape[1:], device=local.device, dtype=local.dtype)
        dist.all_gather_into_tensor(out, local.contiguous())
        return out
    # ragged slices: pad every rank to the largest slice (collectives need equal sizes), trim after
    cmax = max(counts)
    padded = torch.zeros(cmax, *local.shape[1:], device=local.device, dtype=local.dtype)
    padded[: local.shape[0]] = local
    out = torch.empty(cmax * world, *local.shape[1:], device=local.device, dtype=local.dtype)
    dist.all_gather_into_tensor(out, padded)
    return torch.cat([out[r * cmax: r * cmax + c] for r, c in enumerate(counts)])


def allreduce_mean_grads(params: Iterable[torch.nn.Parameter], world: int) -> None:
    """One all-reduce over a flat buffer of every present gradient; result is the mean over ranks
    (each rank's loss is a mean over its own equally sized shard)."""
    if world == 1:
        return
    import torch.distributed as dist

    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat)
    flat.div_(world)
    off = 0
    for g in grads:
        k = g.numel()
        g.copy_(flat[off:off + k].view_as(g))
        off += k


def allreduce_flat_mean(flat: torch.Tensor, world: int) -> None:
    """Mean over ranks of a flat gradient buffer that every p.grad aliases (one collective, in place)."""
    if world == 1:
        return
    import torch.distributed as dist

    dist.all_reduce(flat)
    flat.div_(world)

"""Batch sharding across GPUs: one process per GPU, replicated weights, images split contiguously.

The reference has no multi-device path (no collectives, <= 3 host threads).  Images are
independent units for every operator on the hot path (the reference itself is per-image), so the path shards
with NO data-path collective; the only exchange is gathering the logits (N x 1000 or N x 10 fp32) on rank 0,
done with one NCCL all-gather / gather over NVLink (gloo on CPU in the tests).
"""
from __future__ import annotations

from typing import Callable, List, Optional, Tuple


def shard_bounds(n: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous split of n images over world_size ranks; the first n % world_size ranks get one extra."""
    if world_size <= 0 or not 0 <= rank < world_size:
        raise ValueError("bad rank / world_size")
    base, rem = divmod(n, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_rows(local, group=None, dst: Optional[int] = 0):
    """Concatenate per-rank row blocks [n_r, K] in rank order.  dst=None: every rank gets the result
    (all-gather); otherwise only `dst` does (others get None).  Works for unequal n_r."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    n_local = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
    counts = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(counts, n_local, group=group)
    counts = [int(c.item()) for c in counts]
    k = local.shape[1]
    if len(set(counts)) == 1:
        out = torch.empty((world * counts[0], k), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out if (dst is None or rank == dst) else None
    nmax = max(counts)
    padded = torch.zeros((nmax, k), dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    bufs = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(bufs, padded, group=group)
    if dst is not None and rank != dst:
        return None
    return torch.cat([b[:c] for b, c in zip(bufs, counts)], dim=0)


def run_sharded(run_local: Callable, x_full, group=None, dst: Optional[int] = 0):
    """Split x_full [N, ...] by rank, run `run_local(shard) -> [n_r, K]` on this rank's shard, gather rows.
    `run_local` is Engine.run_torch on GPU ranks; the CPU tests pass a stand-in."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    lo, hi = shard_bounds(x_full.shape[0], world, rank)
    local = run_local(x_full[lo:hi])
    return gather_rows(local, group=group, dst=dst)

"""Harness plumbing for the one-process-per-GPU launch (bench.py, N > 1).

The query path shards as independent units: every rank holds a
replica of the index and a disjoint slice of the batch; there is NO data-path collective.
torch.distributed is used only for the barrier around the timed region, the max-over-ranks
of the timings and (in tests) gathering per-rank results for comparison."""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def env_rank_world() -> tuple[int, int, int]:
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def shard_bounds(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous slice of an n-query batch owned by `rank` (same rule as the C ABI's
    multi-device split, capi.cu slice_for)."""
    return n * rank // world, n * (rank + 1) // world


def max_over_ranks(x: float, device="cpu") -> float:
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return x
    t = torch.tensor([x], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(x: int, device="cpu") -> int:
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return x
    t = torch.tensor([x], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return int(t.item())


def gather_slices(local: torch.Tensor, n: int) -> torch.Tensor | None:
    """Host-side gather of per-rank result slices into one array on rank 0 (tests only)."""
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = [shard_bounds(n, r, world)[1] - shard_bounds(n, r, world)[0] for r in range(world)]
    width = max(sizes)
    pad = torch.zeros(width, dtype=local.dtype)
    pad[: local.numel()] = local
    bufs = [torch.zeros(width, dtype=local.dtype) for _ in range(world)] if rank == 0 else None
    dist.gather(pad, bufs, dst=0)
    if rank != 0:
        return None
    return torch.cat([b[:s] for b, s in zip(bufs, sizes)])

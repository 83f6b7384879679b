"""Multi-GPU plumbing.  Clips are independent, so a batch shards contiguously over ranks with NO
collective in the hot loop; the only exchange is one all-reduce(sum) of six int64 counters
(bit_errors, valid_bits, I_fg, U_fg, I_bg, U_bg) whose ratios give the reference's GLOBAL BER
(scripts/evaluate.py:498-505) and mIoU (scripts/evaluate.py:636-656) - not a mean of per-rank
ratios.  Works with the nccl backend on GPUs and gloo on CPU (tests)."""
from __future__ import annotations

from typing import Tuple

import torch


def shard_range(n_clips: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [start, end) slice of the batch owned by `rank` (sizes differ by at most 1)."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, rem = divmod(n_clips, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def allreduce_counters(counters: torch.Tensor) -> torch.Tensor:
    """Sum the six int64 counters over all ranks (no-op without an initialised process group)."""
    import torch.distributed as dist
    if counters.dtype != torch.int64 or counters.numel() != 6:
        raise ValueError("counters must be int64[6]")
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(counters, op=dist.ReduceOp.SUM)
    return counters

"""Data-parallel sharding of the hot path across the GPUs of one box.

Chunks are independent units (every normalisation on the path is per item; eval-mode BatchNorm is an
affine), so inference shards the batch by item across ranks with NO data-path collective
(SURVEY.md section 8e); torch.distributed is only used to gather results when the caller asks for them and,
in bench.py, for the barrier / max-over-ranks timing.  Host-side logic only: it runs on the gloo backend
in the CPU tests (tests/test_parallel_cpu.py; their gloo workers and stand-ins live in tests/dist_workers.py) and on NCCL on the GPU box.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Tuple

import torch


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced [start, stop) slice of `n_items` for `rank`: the first n % world ranks get one extra."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_sizes(n_items: int, world: int) -> List[int]:
    return [shard_range(n_items, r, world)[1] - shard_range(n_items, r, world)[0] for r in range(world)]


def run_sharded(fn: Callable[[torch.Tensor], torch.Tensor], x: torch.Tensor, gather: bool = True,
                group: Optional["torch.distributed.ProcessGroup"] = None) -> torch.Tensor:
    """Apply the per-item function `fn` (e.g. `model.sample`) to this rank's slice of the global batch `x`
    (every rank passes the same global batch, or at least a tensor of the global batch size) and, if `gather`,
    return the re-assembled global output on every rank.  No collective touches the data path of `fn`."""
    import torch.distributed as dist

    if not dist.is_available() or not dist.is_initialized():
        return fn(x)
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    lo, hi = shard_range(x.shape[0], rank, world)
    local = fn(x[lo:hi]) if hi > lo else None
    if not gather:
        return local
    # ragged all-gather: exchange per-item trailing shape first, then pad-free gather via all_gather_object-free lists
    sizes = shard_sizes(x.shape[0], world)
    tail = torch.tensor(list(local.shape[1:]) if local is not None else [], dtype=torch.long, device="cpu")
    ndim = torch.tensor([tail.numel()], dtype=torch.long)
    nd_all = [torch.zeros_like(ndim) for _ in range(world)]
    dist.all_gather(nd_all, ndim, group=group)
    nd = max(int(t.item()) for t in nd_all)
    tail_pad = torch.zeros(nd, dtype=torch.long)
    tail_pad[: tail.numel()] = tail
    tails = [torch.zeros(nd, dtype=torch.long) for _ in range(world)]
    dist.all_gather(tails, tail_pad, group=group)
    shape_tail = next(tuple(int(v) for v in t.tolist()) for t, s in zip(tails, sizes) if s > 0)
    ref = local if local is not None else x
    outs = [torch.empty((s,) + shape_tail, dtype=ref.dtype if local is not None else torch.float32, device=ref.device) for s in sizes]
    if local is None:
        local = outs[rank]
    pieces = []
    for r in range(world):  # broadcast-based gather keeps ragged shard sizes simple and exact
        buf = local.contiguous() if r == rank else outs[r]
        if sizes[r] > 0:
            dist.broadcast(buf, src=r, group=group)
        pieces.append(buf)
    return torch.cat(pieces, 0)

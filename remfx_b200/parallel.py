"""Data-parallel sharding of the hot path across the GPUs of one box.

Chunks are independent units (every normalisation on the path is per item; eval-mode BatchNorm is an
affine), so inference shards the batch by item across ranks with NO data-path collective
(SURVEY.md section 8e); torch.distributed is only used to gather results when the caller asks for them and,
in bench.py, for the barrier / max-over-ranks timing.  Host-side logic only: it runs on the gloo backend
in the CPU tests (tests/test_parallel_cpu.py) and on NCCL on the GPU box.
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


def _demo_item_op(x: torch.Tensor) -> torch.Tensor:
    """A per-item stand-in for `model.sample` used by the gloo self-test (items independent, like the real path)."""
    y = torch.cumsum(x, dim=-1)
    return (y - y.mean(dim=-1, keepdim=True)) / y.std(dim=-1, keepdim=True).clamp_min(1e-6)


def _gloo_selftest_worker(rank: int, world: int, port: int, n_items: int, q) -> None:
    """Entry point of the world_size>1 CPU test (tests/test_parallel_cpu.py); lives here so spawned workers can import it."""
    import os

    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        x = torch.randn(n_items, 1, 4096, generator=g)
        calls = []

        def fn(xs):
            calls.append(xs.shape[0])
            return _demo_item_op(xs)

        full = run_sharded(fn, x, gather=True)
        ref = _demo_item_op(x)
        lo, hi = shard_range(n_items, rank, world)
        ok = full.shape == ref.shape and torch.allclose(full, ref, rtol=1e-5, atol=1e-6) and sum(calls) == hi - lo
        local = run_sharded(fn, x, gather=False)
        ok = ok and (local is None if hi == lo else torch.allclose(local, ref[lo:hi], rtol=1e-5, atol=1e-6))
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def _gloo_optim_worker(rank: int, world: int, port: int, q) -> None:
    """world_size>1 CPU test of the L5 host logic (tests/test_optim_cpu.py): every rank holds different gradients in its
    flat bucket; after `sync_grads` the bucket holds the SUM and the returned scale turns it into the DDP mean."""
    import os

    import torch.distributed as dist

    from .optim import FlatBucket, sync_grads

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        net = torch.nn.Sequential(torch.nn.Linear(7, 5), torch.nn.Linear(5, 3))
        bucket = FlatBucket(net.parameters())
        x = torch.randn(4, 7, generator=torch.Generator().manual_seed(100 + rank))
        net(x).square().sum().backward()   # autograd accumulates straight into the bucket views
        local = [p.grad.clone() for p in net.parameters()]
        gathered = [[torch.zeros_like(g) for _ in range(world)] for g in local]
        for g, outs in zip(local, gathered):
            dist.all_gather(outs, g)
        scale = sync_grads(bucket.grad)
        ok = abs(scale - 1.0 / world) < 1e-12
        for p, outs in zip(net.parameters(), gathered):
            ok = ok and torch.allclose(p.grad * scale, sum(outs) / world, rtol=1e-6, atol=1e-7)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


class _StubNet(torch.nn.Module):
    """CPU stand-in for a network wrapper in the host-logic tests: `forward((x, y)) -> (loss, out)` with a per-item-mean loss."""

    def __init__(self):
        super().__init__()
        torch.manual_seed(3)
        self.conv = torch.nn.Conv1d(1, 1, 5, padding=2)

    def forward(self, batch):
        x, y = batch
        out = self.conv(x)
        return (out - y).square().mean(), out


class _BucketSGD:
    """CPU stand-in with FusedAdamW's structure (flat bucket, `sync_grads` inside `step`) for the gloo tests."""

    def __init__(self, params, lr: float, group=None):
        from .optim import FlatBucket

        self.bucket, self.lr, self.group = FlatBucket(params), lr, group

    def zero_grad(self):
        self.bucket.zero_grad()

    def step(self):
        from .optim import sync_grads

        self.bucket.collect_grads()
        scale = sync_grads(self.bucket.grad, self.group)
        with torch.no_grad():
            self.bucket.param.add_(self.bucket.grad, alpha=-self.lr * scale)


def _stub_metrics(monkeypatch_target) -> None:
    """Point remfx_b200.train's metric kernels at torch-CPU functions (host-logic tests only; never used by the product)."""
    monkeypatch_target.sisdr_loss = lambda a, b: -(a * b).mean()
    monkeypatch_target.mrstft_loss = lambda a, b: (a - b).abs().mean()


def _gloo_train_worker(rank: int, world: int, port: int, q) -> None:
    """world_size>1 CPU test of the L4/L5 host logic in remfx_b200.train (tests/test_train_cpu.py): every rank runs
    `RemFX.fit_step` on its shard of a global batch; parameters must stay identical across ranks and equal to a single-process
    step on the whole batch, and `sync_dist` metrics must be the mean over ranks."""
    import os

    import torch.distributed as dist

    from . import train as T

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        _stub_metrics(T)
        g = torch.Generator().manual_seed(7)
        x, y = torch.randn(4 * world, 1, 64, generator=g), torch.randn(4 * world, 1, 64, generator=g)
        lo, hi = shard_range(x.shape[0], rank, world)
        mod = T.RemFX(1e-4, 0.95, 0.999, 1e-6, 1e-3, 48000, _StubNet(), max_steps=10)
        opt = _BucketSGD(mod.model.parameters(), lr=0.1)
        for _ in range(3):
            mod.fit_step((x[lo:hi], y[lo:hi], None, None), optimizer=opt)
        # single-process reference on the whole batch
        ref = _StubNet()
        ropt = torch.optim.SGD(ref.parameters(), lr=0.1)
        for _ in range(3):
            ropt.zero_grad()
            ref((x, y))[0].backward()
            ropt.step()
        ok = all(torch.allclose(a, b, rtol=1e-5, atol=1e-6) for a, b in zip(mod.model.parameters(), ref.parameters()))
        # sync_dist metric = mean over ranks of the per-shard values of the LAST step (taken before that step's update)
        local = mod.logged["Input_STFT"].clone()
        want = (x - y).abs().mean()  # equal shard sizes: mean of shard means
        ok = ok and abs(float(local) - float(want)) < 1e-6 and mod.global_step == 3
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()

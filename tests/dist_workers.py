"""Workers and CPU stand-ins of the world_size>1 gloo tests (test_parallel_cpu / test_optim_cpu / test_train_cpu).

Test scaffolding only: spawned processes import this module by name (the tests put the repo root and tests/ on PYTHONPATH).
"""
from __future__ import annotations

import torch

from remfx_b200.parallel import run_sharded, shard_range


def _demo_item_op(x: torch.Tensor) -> torch.Tensor:
    """A per-item stand-in for `model.sample` used by the gloo self-test (items independent, like the real path)."""
    y = torch.cumsum(x, dim=-1)
    return (y - y.mean(dim=-1, keepdim=True)) / y.std(dim=-1, keepdim=True).clamp_min(1e-6)


def _gloo_selftest_worker(rank: int, world: int, port: int, n_items: int, q) -> None:
    """Entry point of the world_size>1 CPU test (tests/test_parallel_cpu.py)."""
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

    from remfx_b200.optim import FlatBucket, sync_grads

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
        from remfx_b200.optim import FlatBucket

        self.bucket, self.lr, self.group = FlatBucket(params), lr, group

    def zero_grad(self, set_to_none: bool = False):
        self.bucket.zero_grad(set_to_none)

    def step(self):
        from remfx_b200.optim import sync_grads

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

    from remfx_b200 import train as T

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

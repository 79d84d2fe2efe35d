"""`not gpu`: host logic of the L5 optimiser step -- flat bucket aliasing, LR schedule, gradient all-reduce on gloo."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

from remfx_b200._lib import RfxError
from remfx_b200.optim import FlatBucket, FusedAdamW, alloc_param_grads, configure_optimizers, multistep_lr
from dist_workers import _gloo_optim_worker


def _net():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Linear(7, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3))


def test_flat_bucket_aliases_params_and_grads():
    net = _net()
    before = [p.detach().clone() for p in net.parameters()]
    b = FlatBucket(net.parameters())
    assert b.numel % 64 == 0 and all(o % 64 == 0 for o in b.offsets)
    for p, o, ref in zip(b.params, b.offsets, before):
        assert torch.equal(p.data, ref)
        assert p.data.data_ptr() == b.param.data_ptr() + 4 * o
        assert p.grad.data_ptr() == b.grad.data_ptr() + 4 * o
    net(torch.randn(4, 7)).sum().backward()          # autograd accumulates in place into the bucket
    assert float(b.grad.abs().sum()) > 0
    for i, p in enumerate(b.params):
        assert torch.equal(p.grad, b.grad_view(i))
    # padding stays zero; gradients re-created by set_to_none are collected back
    mask = torch.ones(b.numel, dtype=torch.bool)
    for p, o in zip(b.params, b.offsets):
        mask[o:o + p.numel()] = False
    assert float(b.grad[mask].abs().sum()) == 0 and float(b.param[mask].abs().sum()) == 0
    for p in b.params:
        p.grad = None
    net(torch.ones(2, 7)).sum().backward()
    want = [p.grad.clone() for p in b.params]
    b.collect_grads()
    for i, (p, w) in enumerate(zip(b.params, want)):
        assert torch.equal(b.grad_view(i), w) and p.grad.data_ptr() == b.grad_view(i).data_ptr()
    b.zero_grad()
    assert float(b.grad.abs().sum()) == 0


class _SinkFn(torch.autograd.Function):
    """A hand-written backward in the style of the network wrappers: gradients written into buffers from alloc_param_grads."""

    @staticmethod
    def forward(ctx, x, *params):
        ctx.save_for_backward(x, *params)
        return sum((p * p).sum() for p in params) + 0.0 * x.sum()

    @staticmethod
    def backward(ctx, d):
        x, *params = ctx.saved_tensors
        grads, zeroed = alloc_param_grads(params)
        _SinkFn.zeroed = zeroed
        for g, p in zip(grads, params):
            if not zeroed:
                g.zero_()
            g.add_(2.0 * p.detach() * d)
        return (None, *[g if p.requires_grad else None for g, p in zip(grads, params)])


def test_gradient_sink_is_adopted_without_copies():
    net = _net()
    ps = list(net.parameters())
    b = FlatBucket(ps)
    x = torch.ones(3)
    # (1) gradients dropped -> one flat zeroed buffer in the bucket's layout, stored by autograd as is, adopted by collect_grads
    b.zero_grad(set_to_none=True)
    assert all(p.grad is None for p in ps)
    _SinkFn.apply(x, *ps).backward()
    assert _SinkFn.zeroed and b._incoming is not None
    flat = b._incoming
    for p, o in zip(ps, b.offsets):
        assert p.grad.data_ptr() == flat.data_ptr() + 4 * o       # autograd kept the view (no clone, no accumulate kernel)
    assert b.collect_grads() == [] and b.grad.data_ptr() == flat.data_ptr()
    for i, p in enumerate(ps):
        assert torch.equal(b.grad_view(i), 2.0 * p.detach()) and p.grad.data_ptr() == b.grad_view(i).data_ptr()
    mask = torch.ones(b.numel, dtype=torch.bool)
    for p, o in zip(ps, b.offsets):
        mask[o:o + p.numel()] = False
    assert float(b.grad[mask].abs().sum()) == 0                   # padding stays zero: whole-bucket kernels remain safe
    # (2) gradients kept as bucket views (the default zero_grad): no sink, autograd accumulates in place as before
    b.zero_grad()
    _SinkFn.apply(x, *ps).backward()
    assert not _SinkFn.zeroed and b.collect_grads() == []
    for i, p in enumerate(ps):
        assert torch.equal(b.grad_view(i), 2.0 * p.detach())
    # (3) a parameter without a gradient this step is reported missing and its slice is zero; a frozen one is skipped
    b.zero_grad(set_to_none=True)
    _SinkFn.apply(x, *ps).backward()
    ps[1].grad = None
    assert b.collect_grads() == [1] and float(b.grad_view(1).abs().sum()) == 0
    # (4) accumulation over two backward calls falls back to the copy path and still sums
    b.zero_grad(set_to_none=True)
    _SinkFn.apply(x, *ps).backward()
    _SinkFn.apply(x, *ps).backward()
    assert b.collect_grads() == []
    for i, p in enumerate(ps):
        assert torch.allclose(b.grad_view(i), 4.0 * p.detach())


def test_multistep_lr_matches_torch_schedule():
    max_steps = 200
    net = _net()
    opt = torch.optim.AdamW(net.parameters(), lr=1e-4)
    sched = torch.optim.lr_scheduler.MultiStepLR(opt, [0.8 * max_steps, 0.95 * max_steps], gamma=0.1)
    for step in range(max_steps):
        assert opt.param_groups[0]["lr"] == pytest.approx(multistep_lr(step, max_steps), rel=1e-12)
        opt.step()
        sched.step()


def test_configure_optimizers_layout_and_no_cpu_fallback():
    net = _net()
    cfg = configure_optimizers(net, max_steps=100)
    opt = cfg["optimizer"]
    assert isinstance(opt, FusedAdamW) and cfg["lr_scheduler"]["interval"] == "step"
    g = opt.param_groups[0]
    assert (g["lr"], g["betas"], g["eps"], g["weight_decay"]) == (1e-4, (0.95, 0.999), 1e-6, 1e-3) and opt.max_grad_norm == 10.0
    sd = opt.state_dict()
    ref = torch.optim.AdamW(_net().parameters())
    assert set(sd) == set(ref.state_dict()) and sd["param_groups"][0]["params"] == ref.state_dict()["param_groups"][0]["params"]
    net(torch.randn(4, 7)).sum().backward()
    with pytest.raises(RfxError):
        opt.step()                                   # CPU tensors: the product path refuses, it never falls back


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_sync_grads_gloo_world2(monkeypatch):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    monkeypatch.setenv("PYTHONPATH", root + os.pathsep + os.path.join(root, "tests") + os.pathsep + os.environ.get("PYTHONPATH", ""))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_optim_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert sorted(q.get(timeout=5) for _ in range(2)) == [(0, True), (1, True)]

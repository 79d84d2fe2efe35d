"""gpu: L5 optimiser kernels (csrc/optim.cu) against torch.optim.AdamW + clip_grad_norm_ (the reference's optimiser,
remfx/models.py:185-191, cfg/config.yaml:119) run in fp32 on the CPU.  Tolerance: 2e-6 relative on parameters after 6 steps."""
import pytest
import torch

from remfx_b200.optim import FusedAdamW

pytestmark = pytest.mark.gpu


def _net(seed=0):
    torch.manual_seed(seed)
    return torch.nn.Sequential(torch.nn.Linear(130, 257), torch.nn.Tanh(), torch.nn.Linear(257, 33), torch.nn.Tanh(), torch.nn.Linear(33, 3))


@pytest.mark.parametrize("clip,gscale", [(10.0, 1.0), (0.05, 1.0), (None, 1.0), (0.5, 30.0)])
def test_adamw_clip_matches_torch(clip, gscale):
    ref = _net()
    ours = _net().cuda()
    kw = dict(lr=1e-3, betas=(0.95, 0.999), eps=1e-6, weight_decay=1e-3)
    opt_ref = torch.optim.AdamW(ref.parameters(), **kw)
    opt = FusedAdamW(ours.parameters(), max_grad_norm=clip, **kw)
    sched_ref = torch.optim.lr_scheduler.MultiStepLR(opt_ref, [3, 5], gamma=0.1)
    sched = torch.optim.lr_scheduler.MultiStepLR(opt, [3, 5], gamma=0.1)
    g = torch.Generator().manual_seed(1)
    for step in range(6):
        x = torch.randn(16, 130, generator=g)
        opt_ref.zero_grad()
        opt.zero_grad()
        (ref(x).square().sum() * gscale).backward()
        (ours(x.cuda()).square().sum() * gscale).backward()
        if clip:
            norm_ref = torch.nn.utils.clip_grad_norm_(ref.parameters(), clip)
        opt_ref.step()
        opt.step()
        sched_ref.step()
        sched.step()
        if clip:
            assert float(opt.total_norm) == pytest.approx(float(norm_ref), rel=2e-5)
        for a, b in zip(ref.parameters(), ours.parameters()):
            err = (a.detach() - b.detach().cpu()).norm() / a.detach().norm()
            assert float(err) < 2e-6, (step, float(err))
    sd, sd_ref = opt.state_dict(), opt_ref.state_dict()
    for i in sd_ref["state"]:
        for k in ("exp_avg", "exp_avg_sq"):
            a, b = sd_ref["state"][i][k], sd["state"][i][k].cpu()
            assert float((a - b).norm() / (a.norm() + 1e-30)) < 1e-4
        assert float(sd["state"][i]["step"]) == float(sd_ref["state"][i]["step"])


def test_grad_sumsq_large_unaligned_tail():
    from remfx_b200 import _lib
    L = _lib.lib()
    n = 33_554_432 + 3
    g = torch.randn(n, device="cuda")
    ws = torch.zeros(32, dtype=torch.float64, device="cuda")
    _lib.check(L.rfx_grad_sumsq(g.data_ptr(), n, ws.data_ptr(), 0, _lib.cur_stream()))
    want = float(g.double().square().sum())
    assert float(ws[0]) == pytest.approx(want, rel=1e-9)
    _lib.check(L.rfx_grad_sumsq(g.data_ptr(), n, ws.data_ptr(), 1, _lib.cur_stream()))
    assert float(ws[0]) == pytest.approx(2 * want, rel=1e-9)


def test_frozen_parameters_are_left_alone_and_moved_storage_is_rebound():
    """torch.optim.AdamW skips parameters without a gradient (no decay, no moments); and a module cast / moved after
    the optimiser was built must keep training (the bucket re-binds the live parameters) instead of silently freezing."""
    ref = _net()
    ours = _net().cuda()
    for net in (ref, ours):
        net[2].weight.requires_grad_(False)   # frozen fine-tuning
    kw = dict(lr=1e-3, betas=(0.95, 0.999), eps=1e-6, weight_decay=1e-2)
    opt_ref = torch.optim.AdamW(ref.parameters(), **kw)
    opt = FusedAdamW(ours.parameters(), max_grad_norm=None, **kw)
    frozen0 = ours[2].weight.detach().clone()
    g = torch.Generator().manual_seed(2)
    for step in range(4):
        x = torch.randn(8, 130, generator=g)
        opt_ref.zero_grad()
        opt.zero_grad()
        ref(x).square().sum().backward()
        ours(x.cuda()).square().sum().backward()
        if step == 2:  # storage leaves the bucket (what module.float() / .to() do to p.data)
            for p in ours.parameters():
                p.data = p.data.clone()
        opt_ref.step()
        opt.step()
        for a, b in zip(ref.parameters(), ours.parameters()):
            err = (a.detach() - b.detach().cpu()).norm() / a.detach().norm()
            assert float(err) < 2e-6, (step, float(err))
    assert torch.equal(ours[2].weight.detach(), frozen0)
    # the live parameters alias the bucket again
    b = opt.bucket
    assert all(p.data_ptr() == b.param_view(i).data_ptr() for i, p in enumerate(b.params))

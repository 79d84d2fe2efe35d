"""`not gpu`: host logic of the training-step mirror (remfx_b200.train.RemFX, rows L4 / L5) with CPU stand-ins for the
network and the metric kernels -- logged names, step sequencing, scheduler stepping, data-parallel averaging on gloo."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

from remfx_b200 import train as T
from remfx_b200.optim import FusedAdamW
from dist_workers import _gloo_train_worker, _stub_metrics, _StubNet


def _batch(B=3, n=64, seed=1):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(B, 1, n, generator=g), torch.randn(B, 1, n, generator=g), None, None


@pytest.fixture()
def stubbed(monkeypatch):
    monkeypatch.setattr(T, "sisdr_loss", T.sisdr_loss)
    monkeypatch.setattr(T, "mrstft_loss", T.mrstft_loss)
    _stub_metrics(T)
    yield


def _module(**kw):
    return T.RemFX(1e-4, 0.95, 0.999, 1e-6, 1e-3, 48000, _StubNet(), **kw)


@pytest.mark.parametrize("mode,method", [("train", "training_step"), ("valid", "validation_step"), ("test", "test_step")])
def test_common_step_logs_reference_names(stubbed, mode, method):
    m = _module()
    x, y, _, _ = b = _batch()
    loss = getattr(m, method)(b, 0)
    assert loss.requires_grad and loss.dim() == 0
    assert set(m.logged) == {f"{mode}_loss", f"{mode}_SISDR", f"{mode}_STFT", "Input_SISDR", "Input_STFT"}
    with torch.no_grad():
        out = m.model((x, y))[1]
    assert float(m.logged[f"{mode}_loss"]) == pytest.approx(float(loss.detach()))
    assert float(m.logged[f"{mode}_SISDR"]) == pytest.approx(float((out * y).mean()))   # negated stub "loss"
    assert float(m.logged["Input_STFT"]) == pytest.approx(float((x - y).abs().mean()))
    assert not any(v.requires_grad for v in m.logged.values())


def test_metric_block_can_be_switched_off(stubbed):
    m = _module()
    m.compute_metrics = False
    m.training_step(_batch(), 0)
    assert set(m.logged) == {"train_loss"}


def test_target_is_causal_cropped_to_the_output(stubbed):
    class Short(_StubNet):
        def forward(self, batch):
            x, y = batch
            out = self.conv(x)[..., :-7]
            return out.square().mean(), out

    seen = {}
    T.mrstft_loss = lambda a, b: seen.setdefault((a.shape[-1], b.shape[-1]), (a - b).abs().mean())
    m = T.RemFX(1e-4, 0.95, 0.999, 1e-6, 1e-3, 48000, Short())
    x, y, _, _ = b = _batch()
    m.training_step(b, 0)
    assert (57, 57) in seen and (64, 64) in seen
    # causal_crop keeps [L-1-l, L-1): the reference's off-by-one (remfx/utils.py:208-211)
    from remfx_b200.ops import causal_crop

    assert torch.equal(causal_crop(y, 57), y[..., 6:63])


def test_fit_step_runs_the_trainer_sequence(stubbed):
    m = _module(max_steps=5)
    opt = torch.optim.AdamW(m.model.parameters(), lr=1e-2)
    sched = torch.optim.lr_scheduler.MultiStepLR(opt, [0.8 * 5, 0.95 * 5], gamma=0.1)
    before = [p.detach().clone() for p in m.model.parameters()]
    losses = [float(m.fit_step(_batch(), i, optimizer=opt, scheduler=sched)) for i in range(5)]
    assert m.global_step == 5 and losses[-1] < losses[0]
    assert all(not torch.equal(a, p) for a, p in zip(before, m.model.parameters()))
    # the scheduler was stepped once per batch: milestone 4.0 is hit after the 4th step (4.75 never equals an integer epoch,
    # exactly as torch's MultiStepLR treats the reference's float milestones)
    assert opt.param_groups[0]["lr"] == pytest.approx(1e-3)


def test_configure_optimizers_structure_and_max_steps_source(stubbed):
    m = _module()
    with pytest.raises(ValueError):
        m.configure_optimizers()
    m.trainer = type("Trainer", (), {"max_steps": 50})()   # what Lightning attaches (cfg/config.yaml:113)
    cfg = m.configure_optimizers()
    assert isinstance(cfg["optimizer"], FusedAdamW) and cfg["optimizer"].max_grad_norm == 10.0
    assert cfg["lr_scheduler"]["interval"] == "step" and cfg["lr_scheduler"]["scheduler"].milestones == {40.0: 1, 47.5: 1}
    g = cfg["optimizer"].param_groups[0]
    assert (g["lr"], g["betas"], g["eps"], g["weight_decay"]) == (1e-4, (0.95, 0.999), 1e-6, 1e-3)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_data_parallel_fit_step_gloo_world2(monkeypatch):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    monkeypatch.setenv("PYTHONPATH", root + os.pathsep + os.path.join(root, "tests") + os.pathsep + os.environ.get("PYTHONPATH", ""))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_train_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert sorted(q.get(timeout=5) for _ in range(2)) == [(0, True), (1, True)]


def test_common_step_matches_the_unchanged_reference_harness(monkeypatch):
    """Live against `remfx.models.RemFX.common_step` (remfx/models.py:217-256) run through the shim (skipped where
    /root/reference is absent): same stub network, the metric classes both sides use are the oracle restatements of auraloss --
    logged names, the SISDR negation, the causal crop of the target and the returned loss must agree."""
    from oracle import loss as oloss
    from oracle import refshim

    if not refshim.available():
        pytest.skip("/root/reference not present (GPU box)")
    R = refshim.ref_modules()

    class Short(_StubNet):   # output shorter than the target: exercises the crop
        def forward(self, batch):
            x, y = batch
            out = self.conv(x)[..., 5:-4]
            return (out - y[..., 8:-1]).abs().mean(), out

    ref = R.models.RemFX(lr=1e-4, lr_beta1=0.95, lr_beta2=0.999, lr_eps=1e-6, lr_weight_decay=1e-3, sample_rate=48000, network=Short())
    ref_logged = {}
    ref.log = lambda name, value, *a, **k: ref_logged.__setitem__(name, float(value.detach()))
    monkeypatch.setattr(T, "sisdr_loss", oloss.sisdr_loss)
    monkeypatch.setattr(T, "mrstft_loss", oloss.mrstft)
    mine = T.RemFX(1e-4, 0.95, 0.999, 1e-6, 1e-3, 48000, Short())
    g = torch.Generator().manual_seed(11)
    x, y = 0.1 * torch.randn(2, 1, 6000, generator=g), 0.1 * torch.randn(2, 1, 6000, generator=g)
    for mode, method in (("train", "training_step"), ("valid", "validation_step"), ("test", "test_step")):
        ref_logged.clear()
        mine.logged.clear()
        rl = getattr(ref, method)((x, y, None, None), 0)
        ml = getattr(mine, method)((x, y, None, None), 0)
        assert float(ml.detach()) == pytest.approx(float(rl.detach()), rel=1e-6)
        assert set(mine.logged) == set(ref_logged) == {f"{mode}_loss", f"{mode}_SISDR", f"{mode}_STFT", "Input_SISDR", "Input_STFT"}
        for k, v in ref_logged.items():
            assert float(mine.logged[k]) == pytest.approx(v, rel=1e-6), (mode, k)
    # configure_optimizers (remfx/models.py:185-206): same hyper-parameters, schedule and return structure
    ref.trainer = type("Trainer", (), {"max_steps": 50})()
    mine.trainer = ref.trainer
    rc, mc = ref.configure_optimizers(), mine.configure_optimizers()
    assert set(rc) == set(mc) and {k: v for k, v in rc["lr_scheduler"].items() if k != "scheduler"} == \
        {k: v for k, v in mc["lr_scheduler"].items() if k != "scheduler"}
    rg, mg = rc["optimizer"].param_groups[0], mc["optimizer"].param_groups[0]
    for k in ("lr", "betas", "eps", "weight_decay"):
        assert tuple(rg[k]) == tuple(mg[k]) if k == "betas" else rg[k] == mg[k], k
    rs, ms = rc["lr_scheduler"]["scheduler"], mc["lr_scheduler"]["scheduler"]
    assert type(rs) is type(ms) and rs.milestones == ms.milestones and rs.gamma == ms.gamma
    assert len(rg["params"]) == len(mg["params"]) == len(list(Short().parameters()))

"""-m gpu: fused MR-STFT + L1 loss kernels vs the oracle restatement of auraloss (parity unpinned upstream)."""
import math

import pytest
import torch

from oracle import loss as oloss
from oracle import stft as ostft
from oracle import weights

pytestmark = pytest.mark.gpu


def _terms(a, b):
    from remfx_b200.losses import remfx_loss_terms

    return remfx_loss_terms(a.cuda(), b.cuda()).cpu()


@pytest.mark.parametrize("B,T", [(1, 16384), (3, 40000), (2, 262144)])
def test_loss_matches_oracle(B, T):
    a = weights.synth_audio(B * 7 + 1, B, T)
    b = weights.synth_audio(B * 7 + 2, B, T) * 0.7 + 0.3 * a
    res = _terms(a, b)
    ref = oloss.remfx_loss(a, b)
    assert abs(float(res[0]) - float(ref)) < 1e-4 * abs(float(ref)), (float(res[0]), float(ref))
    assert abs(float(res[1]) - float(oloss.mrstft(a, b))) < 1e-4 * float(oloss.mrstft(a, b))
    assert abs(float(res[2]) - float((a - b).abs().mean())) < 1e-5 * float((a - b).abs().mean())


def test_loss_anchors():
    """Size-independent properties: L(a, a) = 0 and MRSTFT(0.5 a, a) = 0.5 + ln 2."""
    a = weights.synth_audio(3, 2, 65536)
    assert float(_terms(a, a)[0]) == 0.0
    assert abs(float(_terms(0.5 * a, a)[1]) - (0.5 + math.log(2.0))) < 1e-4


def test_loss_on_cropped_target_view():
    """TCN path: target = causal_crop(target, len(out)) is passed as a strided view (remfx/models.py:383-385)."""
    from remfx_b200.losses import remfx_loss
    from remfx_b200.ops import causal_crop

    out = weights.synth_audio(5, 2, 20000)
    tgt = weights.synth_audio(6, 2, 32277)
    got = remfx_loss(out.cuda(), causal_crop(tgt.cuda(), 20000))
    ref = oloss.remfx_loss(out, ostft.causal_crop(tgt, 20000))
    assert abs(float(got) - float(ref)) < 1e-4 * abs(float(ref))


def test_umx_forward_returns_loss_and_output():
    from remfx_b200.models import OpenUnmixModel
    from oracle import umx as oumx

    sd = weights.umx_state(7)
    m = OpenUnmixModel(sample_rate=48000)
    m.load_state_dict(sd)
    m = m.cuda().eval()
    x, t = weights.synth_audio(8, 2, 32768), weights.synth_audio(9, 2, 32768)
    loss, out = m((x.cuda(), t.cuda()))
    rl, ro = oumx.forward((x, t), sd)
    assert loss.dim() == 0 and out.shape == (2, 1, 32768)
    assert abs(float(loss) - float(rl)) < 1e-4 * abs(float(rl))


@pytest.mark.parametrize("B,T", [(1, 8192), (3, 20000), (2, 65536)])
def test_loss_backward_matches_autograd_of_the_oracle(B, T):
    """d loss / d out from rfx_remfx_loss_backward vs torch autograd through the oracle restatement (CPU, fp64 graph)."""
    from remfx_b200.losses import remfx_loss

    a = weights.synth_audio(B * 11 + 1, B, T)
    b = weights.synth_audio(B * 11 + 2, B, T) * 0.7 + 0.3 * a
    xg = a.clone().cuda().requires_grad_(True)
    loss = remfx_loss(xg, b.cuda())
    (2.5 * loss).backward()   # an upstream factor, to check grad_loss is honoured
    g = xg.grad.cpu()
    xr = a.clone().double().requires_grad_(True)
    ref_loss = oloss.remfx_loss(xr, b.double())
    (2.5 * ref_loss).backward()
    gr = xr.grad
    assert abs(float(loss.detach()) - float(ref_loss.detach())) < 1e-4 * abs(float(ref_loss.detach()))
    err = float((g.double() - gr).norm() / gr.norm())
    # the L1 part is +-c exactly; the spectral part carries the fp32 FFT error
    assert err < 1e-4, err
    # spectral part alone (remove the sign term, which dominates the norm)
    l1 = 2.5 * 100.0 / (B * T) * torch.sign(a.double() - b.double()).reshape(gr.shape)
    err_spec = float(((g.double() - l1) - (gr - l1)).norm() / (gr - l1).norm())
    # fp32 spectra against an fp64 graph: the log-magnitude term divides by the magnitude, so the weakest bins set the error
    assert err_spec < 5e-4, err_spec


def test_loss_backward_zero_at_identical_signals_and_no_grad_path():
    from remfx_b200.losses import remfx_loss

    a = weights.synth_audio(5, 2, 16384).cuda()
    with torch.no_grad():
        assert float(remfx_loss(a, a)) == 0.0
    xg = (0.5 * a).clone().requires_grad_(True)
    loss = remfx_loss(xg.view(2, 1, -1), a.view(2, 1, -1))
    loss.backward()
    assert xg.grad.shape == xg.shape and torch.isfinite(xg.grad).all() and float(xg.grad.abs().sum()) > 0

"""-m gpu: properties checked at BASELINE.json's FULL sizes (32 x 262144 @ 48 kHz), where running the CPU oracle on everything
would take too long: round trips, item independence / permutation equivariance, anchors, spot checks against the oracle."""
import math

import pytest
import torch

from oracle import loss as oloss
from oracle import umx as oumx
from oracle import weights
from tests.util import relrms

pytestmark = pytest.mark.gpu

B, T = 32, 262144


def test_stft_istft_round_trip_full_size():
    """umx/tests/test_transforms.py:42-51 at the benchmark shape: RMSE of istft(stft(x)) - x below 1e-6 (all four n_fft kernels)."""
    from remfx_b200 import ops

    x = weights.synth_audio(1, B, T).cuda().view(B, 1, T)
    for n_fft, hop in ((2048, 512), (4096, 1024), (1024, 256), (512, 128)):
        X = ops.stft(x, n_fft=n_fft, n_hop=hop)
        y = ops.istft(X, n_fft=n_fft, n_hop=hop, length=T)
        rmse = float(torch.sqrt(torch.mean((y.reshape(B, T) - x.reshape(B, T)) ** 2)))
        assert rmse < 1e-6, (n_fft, rmse)


def test_umx_full_batch_items_independent_and_spot_checked():
    """Open-Unmix at 32 x 262144: permuting the batch permutes the outputs bit for bit (items never interact: per-item
    arithmetic does not depend on the batch slot), pipeline == sample to 5e-6, and three items are checked against the oracle."""
    from remfx_b200.models import OpenUnmixModel

    sd = weights.umx_state(3)
    m = OpenUnmixModel(n_fft=2048, hop_length=512, n_channels=1, alpha=0.3, sample_rate=48000)
    m.load_state_dict(sd)
    m = m.cuda().eval()
    x = weights.synth_audio(7, B, T)
    xd = x.cuda()
    out = m.sample(xd)
    perm = torch.randperm(B, generator=torch.Generator().manual_seed(0))
    out_p = m.sample(xd[perm.cuda()].contiguous())
    assert torch.equal(out_p, out[perm.cuda()])
    pipe = m.pipeline("cuda:0")
    seqs = [pipe.push(xd), pipe.push(xd[perm.cuda()].contiguous())]
    pipe.flush()
    o1, o2 = pipe.wait(seqs[0]), pipe.wait(seqs[1])
    assert relrms(o1.cpu(), out.cpu()) < 5e-6
    assert torch.equal(o2, o1[perm.cuda()])
    idx = [0, 13, 31]
    ref = oumx.sample(x[idx], sd)
    assert relrms(out[idx].cpu(), ref) < 1e-4
    assert relrms(o1[idx].cpu(), ref) < 1e-4


def test_loss_anchors_and_gradient_sum_full_size():
    """L(a, a) = 0, MRSTFT(a/2, a) = 1/2 + ln 2 at 32 x 262144; the gradient of the L1 term alone sums to the
    closed form, and a central finite difference towards the target matches <grad, direction>."""
    from remfx_b200.losses import remfx_loss, remfx_loss_terms

    a = weights.synth_audio(3, B, T).cuda()
    assert float(remfx_loss_terms(a, a)[0]) == 0.0
    assert abs(float(remfx_loss_terms(0.5 * a, a)[1]) - (0.5 + math.log(2.0))) < 1e-4
    b = weights.synth_audio(4, B, T).cuda()
    xg = (0.7 * a + 0.3 * b).clone().requires_grad_(True)
    loss = remfx_loss(xg, b)
    loss.backward()
    g = xg.grad
    # direction = towards the target (a random direction is nearly orthogonal to the gradient and drowns in the fp32 resolution
    # of the loss value; the gradient direction itself weights the near-silent bins, where log-magnitudes are far from linear)
    d = (b - xg.detach())
    eps = 1e-3
    with torch.no_grad():
        lp = float(remfx_loss_terms(xg + eps * d, b)[0].double())
        lm = float(remfx_loss_terms(xg - eps * d, b)[0].double())
    fd = (lp - lm) / (2 * eps)
    an = float((g.double() * d.double()).sum())
    assert abs(fd - an) < 3e-2 * abs(an) + 1e-4, (fd, an)
    # spot check of three items against the oracle restatement
    idx = [0, 15, 31]
    ref = oloss.remfx_loss(xg.detach()[idx].cpu(), b[idx].cpu())
    val = remfx_loss_terms(xg.detach()[idx].contiguous(), b[idx].contiguous())[0]
    assert abs(float(val) - float(ref)) < 1e-4 * abs(float(ref))

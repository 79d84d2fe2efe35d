"""-m gpu: dense-layer kernels (tcgen05 bf16x3 and fp32 FFMA) and the LSTM recurrence vs fp64 / the oracle."""
import pytest
import torch

from oracle import umx as oumx
from tests.util import relrms

pytestmark = pytest.mark.gpu


def _ops():
    from remfx_b200 import ops

    return ops


def _case(M, N, K, seed):
    g = torch.Generator().manual_seed(seed)
    A = torch.randn(M, K, generator=g)
    W = torch.randn(N, K, generator=g) / K ** 0.5
    s1 = 0.5 + torch.rand(N, generator=g)
    t1 = torch.randn(N, generator=g)
    s2 = 0.5 + torch.rand(N, generator=g)
    t2 = torch.randn(N, generator=g)
    ref = ((A.double() @ W.double().t()) * s1.double() + t1.double()) * s2.double() + t2.double()
    return A, W, s1, t1, s2, t2, ref


@pytest.mark.parametrize("impl", ["simt", "tc"])
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (200, 512, 512), (513, 1025, 512), (300, 512, 1025), (77, 2048, 512)])
def test_linear_affine(impl, M, N, K):
    ops = _ops()
    A, W, s1, t1, s2, t2, ref = _case(M, N, K, M + N + K)
    out = ops.linear(A.cuda(), W.cuda(), s1.cuda(), t1.cuda(), s2.cuda(), t2.cuda(), act=None, impl=impl)
    err = relrms(out, ref)
    assert err < (2e-5 if impl == "tc" else 2e-6), err


@pytest.mark.parametrize("impl", ["simt", "tc"])
@pytest.mark.parametrize("act", ["tanh", "relu", "sigmoid"])
def test_linear_activations(impl, act):
    ops = _ops()
    A, W, s1, t1, _, _, _ = _case(257, 384, 320, 11)
    pre = (A.double() @ W.double().t()) * s1.double() + t1.double()
    ref = {"tanh": torch.tanh, "relu": torch.relu, "sigmoid": torch.sigmoid}[act](pre)
    out = ops.linear(A.cuda(), W.cuda(), s1.cuda(), t1.cuda(), act=act, impl=impl)
    assert relrms(out, ref) < 2e-5


def test_tc_beats_single_pass_bf16_precision():
    """The 3-product split must be far more accurate than one bf16 pass (~4e-3), i.e. fp32-grade."""
    ops = _ops()
    A, W, *_ = _case(256, 256, 1024, 5)
    ref = A.double() @ W.double().t()
    out = ops.linear(A.cuda(), W.cuda(), impl="tc")
    assert relrms(out, ref) < 1e-5


@pytest.mark.parametrize("H,B,F", [(192, 7, 50), (192, 24, 200), (384, 3, 40), (384, 9, 128)])
def test_lstm_layer_other_hidden_sizes(H, B, F):
    """HDemucs BLSTMs: hidden 192 (layer 4) and 384 (layer 5), input size == hidden size (TA:738)."""
    _lstm_case(H, H, B, F, "mma")


@pytest.mark.parametrize("impl", ["mma", "ffma"])
@pytest.mark.parametrize("B,F", [(1, 5), (4, 33), (6, 40), (19, 70)])
def test_lstm_layer_vs_oracle(B, F, impl):
    _lstm_case(256, 512, B, F, impl)


@pytest.mark.parametrize("B,F", [(1, 5), (16, 20), (21, 33), (32, 64), (40, 17)])
def test_lstm_layer_tcgen05(B, F):
    """tcgen05 recurrence (W_hh as the TMEM A operand, 16 slots per cluster) vs the explicit-loop oracle and the mma.sync kernel."""
    out_tc = _lstm_case(256, 512, B, F, "tc")
    out_mma = _lstm_case(256, 512, B, F, "mma", slots=8)
    assert relrms(out_tc, out_mma) < 2e-6
    out_tc32 = _lstm_case(256, 512, B, F, "tc", slots=32)   # 32 slots per cluster (MMA N = 32)
    assert relrms(out_tc32, out_mma) < 2e-6


def _lstm_case(H, I, B, F, impl, slots=0):
    ops = _ops()
    g = torch.Generator().manual_seed(B * 100 + F)
    k = H ** -0.5
    st = {}
    for sfx in ("", "_reverse"):
        st[f"l.weight_ih_l0{sfx}"] = (torch.rand(4 * H, I, generator=g) * 2 - 1) * k
        st[f"l.weight_hh_l0{sfx}"] = (torch.rand(4 * H, H, generator=g) * 2 - 1) * k
        st[f"l.bias_ih_l0{sfx}"] = (torch.rand(4 * H, generator=g) * 2 - 1) * k
        st[f"l.bias_hh_l0{sfx}"] = (torch.rand(4 * H, generator=g) * 2 - 1) * k
    x = torch.randn(F, B, I, generator=g)
    ref = oumx.lstm_explicit(x, st, "l", layers=1)  # (F, B, 2H)
    # input projections (+ both biases), row = b*F + t, column = dir*4H + gate*H + unit
    G = torch.cat([x @ st[f"l.weight_ih_l0{s}"].t() + st[f"l.bias_ih_l0{s}"] + st[f"l.bias_hh_l0{s}"] for s in ("", "_reverse")], -1)
    G = G.permute(1, 0, 2).reshape(B * F, 8 * H).contiguous()
    Whh = torch.stack([st["l.weight_hh_l0"], st["l.weight_hh_l0_reverse"]])
    out = ops.lstm_layer(G.cuda(), Whh.cuda(), B, F, impl=impl, slots=slots)
    out = out.view(B, F, 2 * H).permute(1, 0, 2)
    err = relrms(out, ref)
    print(f"lstm {impl} H={H} B={B} F={F} slots={slots} rel-RMS {err:.3e}")
    assert err < 1e-5, err
    return out.cpu()

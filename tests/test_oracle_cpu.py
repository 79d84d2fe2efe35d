"""`not gpu`: pin the oracle restatements against the golden vectors produced by the UNCHANGED
reference modules (oracle/make_golden.py), against torch.stft/istft (the calls the reference makes),
and -- when /root/reference is present -- against the live reference."""
import numpy as np
import pytest
import torch

from oracle import loss as oloss
from oracle import refshim
from oracle import stft as ostft
from oracle import tcn as otcn
from oracle import umx as oumx
from oracle import weights
from tests.util import golden, relrms


def test_stft_matches_torch_and_golden():
    g = golden("stft_kat.npz")
    x = weights.synth_audio(int(g["seed"]), int(g["B"]), int(g["T"]))[:, 0]
    win = torch.hann_window(2048)
    Z = ostft.stft(x, 2048, 512, win)
    Zt = torch.stft(x, 2048, 512, window=win, return_complex=True)
    assert relrms(torch.view_as_real(Z), torch.view_as_real(Zt)) < 1e-6
    assert relrms(torch.view_as_real(Z), torch.from_numpy(g["Z"])) < 1e-6
    y = ostft.istft(Z, 2048, 512, win, length=x.shape[-1])
    assert relrms(y, torch.from_numpy(g["y"])) < 1e-6


@pytest.mark.parametrize("nfft", [1024, 2048, 4096])
@pytest.mark.parametrize("hopdiv", [2, 4])
@pytest.mark.parametrize("T", [4096, 44100])
def test_stft_istft_roundtrip(nfft, hopdiv, T):
    """umx/tests/test_transforms.py:42-51: round trip RMSE < 1e-6."""
    hop = nfft // hopdiv
    x = torch.rand(2, T, generator=torch.Generator().manual_seed(T + nfft))
    win = torch.hann_window(nfft)
    Z = ostft.stft(x, nfft, hop, win)
    Zt = torch.stft(x, nfft, hop, window=win, return_complex=True)
    assert (Z - Zt).abs().max() < 1e-3 * Zt.abs().max()
    y = ostft.istft(Z, nfft, hop, win, length=T)
    assert float(torch.sqrt(((x - y) ** 2).mean())) < 1e-6


def test_short_window_padding_matches_torch():
    x = weights.synth_audio(5, 1, 8192)[:, 0]
    for n_fft, hop, win in oloss.RESOLUTIONS:
        w = torch.hann_window(win)
        Zt = torch.stft(x, n_fft, hop, win, w, return_complex=True)
        Z = ostft.stft(x, n_fft, hop, ostft.padded_window(win, n_fft))
        assert relrms(torch.view_as_real(Z), torch.view_as_real(Zt)) < 1e-6


def test_mrstft_anchors():
    a = weights.synth_audio(7, 2, 16384)
    assert float(oloss.mrstft(a, a)) == 0.0
    assert abs(float(oloss.mrstft(0.5 * a, a)) - (0.5 + np.log(2.0))) < 1e-5


def test_crops():
    x = torch.arange(10.0)[None]
    assert ostft.center_crop(x, 4).tolist() == [[3.0, 4.0, 5.0, 6.0]]
    assert ostft.causal_crop(x, 4).tolist() == [[5.0, 6.0, 7.0, 8.0]]  # drops the last sample (reference quirk)


def test_umx_oracle_matches_golden():
    g = golden("umx_sample.npz")
    sd = weights.umx_state(int(g["wseed"]))
    assert abs(weights.checksum(sd) - float(g["wsum"])) < 1e-6 * abs(float(g["wsum"]))
    x = weights.synth_audio(int(g["xseed"]), int(g["B"]), int(g["T"]))
    t = weights.synth_audio(int(g["tseed"]), int(g["B"]), int(g["T"]))
    ref = torch.from_numpy(g["out"])
    assert relrms(oumx.sample(x, sd), ref) < 1e-5
    assert relrms(oumx.sample(x, sd, fast_lstm=False, wiener_trig=False), ref) < 1e-5
    loss, _ = oumx.forward((x, t), sd)
    assert abs(float(loss) - float(g["loss"])) < 1e-4 * abs(float(g["loss"]))


def test_tcn_oracle_matches_golden():
    g = golden("tcn_forward.npz")
    sd = weights.tcn_state(int(g["wseed"]))
    assert abs(weights.checksum(sd) - float(g["wsum"])) < 1e-6 * abs(float(g["wsum"]))
    x = weights.synth_audio(int(g["xseed"]), int(g["B"]), int(g["T"]))
    t = weights.synth_audio(int(g["tseed"]), int(g["B"]), int(g["T"]))
    ref = torch.from_numpy(g["out"])
    loss, out = otcn.forward((x, t), sd)
    assert out.shape == ref.shape
    assert relrms(out, ref) < 1e-5
    assert relrms(otcn.sample(x, sd, fused=True), ref) < 1e-5
    assert abs(float(loss) - float(g["loss"])) < 1e-4 * abs(float(g["loss"]))


def test_cnn14_oracle_matches_golden():
    from oracle import cnn14 as ocnn

    g = golden("cnn14_decisions.npz")
    sd = weights.cnn14_state(int(g["wseed"]))
    assert abs(weights.checksum(sd) - float(g["wsum"])) < 1e-6 * abs(float(g["wsum"]))
    x = weights.synth_diverse(int(g["first_xseed"]), 4, int(g["T"]))
    lg = ocnn.logits(x, sd)
    assert (lg - torch.from_numpy(g["logits"][:4])).abs().max() < 1e-3
    assert torch.equal(ocnn.decisions(x, sd), torch.from_numpy(g["decisions"][:4]).long())


@pytest.mark.skipif(not refshim.available(), reason="/root/reference not present (GPU box)")
def test_oracle_matches_live_reference():
    R = refshim.ref_modules()
    sd = weights.umx_state(3)
    m = R.models.OpenUnmixModel(n_fft=2048, hop_length=512, n_channels=1, alpha=0.3, sample_rate=48000)
    m.load_state_dict(sd, strict=True)
    m.eval()
    x = weights.synth_audio(41, 1, 8192)
    with torch.no_grad():
        ref = m.sample(x)
    assert relrms(oumx.sample(x, sd), ref) < 1e-5
    X = R.utils.spectrogram(x, torch.hann_window(2048), 2048, 512, 0.3)
    assert relrms(ostft.spectrogram(x, torch.hann_window(2048), 2048, 512, 0.3), X) < 1e-6


def test_training_loss_decomposes_over_data_parallel_shards():
    """with the per-item spectral-convergence ratio (auraloss >= 0.4, the form restated in oracle/loss.py) the
    training loss of a batch is the mean of its shards' losses, so averaging the shards' gradients (what the one all-reduce
    of a data-parallel step does) reproduces the single-process gradient exactly up to rounding."""
    from oracle import loss as oloss
    from oracle import tcn as otcn

    sd = weights.tcn_state(3, nblocks=3, width=64)
    x = weights.synth_audio(50, 4, 3000)
    y = weights.synth_audio(51, 4, 3000)

    def grads(lo, hi):
        st = {k: v.double().clone().requires_grad_(True) for k, v in sd.items()}
        loss, _ = otcn.forward((x[lo:hi].double(), y[lo:hi].double()), st)
        loss.backward()
        return float(loss.detach()), {k: v.grad for k, v in st.items()}

    full_loss, full = grads(0, 4)
    (la, a), (lb, b) = grads(0, 2), grads(2, 4)
    assert abs(0.5 * (la + lb) - full_loss) < 1e-12 * abs(full_loss)
    for k in full:
        assert torch.allclose(0.5 * (a[k] + b[k]), full[k], rtol=1e-9, atol=1e-12), k
    assert abs(float(oloss.remfx_loss(x, y)) - 0.5 * (float(oloss.remfx_loss(x[:2], y[:2])) + float(oloss.remfx_loss(x[2:], y[2:])))) < 1e-5


def _oracle_tcn_grads(sd, x, r):
    from oracle import tcn as otcn

    st = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    out = otcn.sample(x, st)
    (out * r).sum().backward()
    return out.detach(), {k: v.grad for k, v in st.items()}


def test_tcn_oracle_gradients_match_reference_golden():
    """The gradient oracle of the TCN backward tests = torch autograd through oracle/tcn.py; pinned here against gradients
    the UNCHANGED reference TCNModel produced under autograd (tests/golden/tcn_backward.npz, oracle/make_golden.py)."""
    from tests.util import tcn_backward_case

    sd, x, r, out_ref, g_ref, _ = tcn_backward_case(golden("tcn_backward.npz"))
    out, g = _oracle_tcn_grads(sd, x, r)
    assert relrms(out, out_ref) < 1e-6
    assert set(g) == set(g_ref) and len(g) == 14
    for k in g_ref:
        assert relrms(g[k], g_ref[k]) < 1e-5, k


@pytest.mark.skipif(not refshim.available(), reason="/root/reference not present (GPU box)")
def test_tcn_oracle_gradients_match_live_reference_with_prelu_kinks():
    """Same check with the reference's PReLU slopes, run live: the oracle issues the same torch ops as the reference, so even
    the kink decisions agree and the gradients are equal to rounding."""
    from oracle.make_golden import reference_tcn_gradients

    sd = weights.tcn_state(45, nblocks=4, width=64)
    x = weights.synth_audio(46, 2, 2500)
    Lout = 2500 - sum(6 * 2 ** n for n in range(4))
    r = torch.randn(2, 1, Lout, generator=torch.Generator().manual_seed(47))
    out_ref, g_ref = reference_tcn_gradients(refshim.ref_modules(), sd, x, r, 4, 64)
    out, g = _oracle_tcn_grads(sd, x, r)
    assert relrms(out, out_ref) < 1e-6
    for k in g_ref:
        assert relrms(g[k], g_ref[k]) < 1e-5, k


def test_hdemucs_gradient_oracle_taps():
    """Backward oracle of the Hybrid-Demucs step (oracle/hdemucs.py:grad_taps): every parameter receives a gradient, the
    per-layer activation gradients have the layers' output shapes, and a directional finite difference of the training loss
    agrees with <grad, direction> -- the harness the GPU backward will be checked against layer by layer."""
    from oracle import hdemucs as ohd

    m = ohd.build(seed=2, layerscale=0.1).double()
    T = 16384
    x = weights.synth_audio(70, 1, T).double()
    y = weights.synth_audio(71, 1, T).double()
    names = ["freq_encoder.0", "freq_encoder.4.dconv.layers.0.3", "freq_encoder.4.dconv.layers.0.4", "time_encoder.1", "freq_decoder.5",
             "time_decoder.4"]
    r = ohd.grad_taps(x, y, m, names)
    assert r["output"].shape == (1, 1, T) and r["loss"].dim() == 0
    # the "empty" last time-encoder layer only runs its conv (TA:163-164): its norm1 parameters are never used
    assert {k for k, _ in m.named_parameters()} - set(r["param_grads"]) == {"time_encoder.4.norm1.weight", "time_encoder.4.norm1.bias"}
    assert set(r["act_grads"]) == set(names)
    fwd = ohd.taps(x, m, names)
    for n in names:
        assert r["act_grads"][n].shape == fwd[n].shape, n
    # directional derivative along a seeded direction restricted to a few tensors of each kind
    g = torch.Generator().manual_seed(9)
    keys = ["freq_encoder.0.conv.weight", "freq_encoder.4.dconv.layers.0.3.lstm.weight_hh_l0", "freq_encoder.5.dconv.layers.1.4.query_decay.weight",
            "time_decoder.2.conv_tr.weight", "freq_decoder.1.norm1.weight", "freq_emb.embedding.weight"]
    params = dict(m.named_parameters())
    dirs = {k: torch.randn(params[k].shape, generator=g, dtype=torch.float64) for k in keys}
    from oracle import loss as oloss

    def loss_at(eps):
        with torch.no_grad():
            for k in keys:
                params[k].add_(eps * dirs[k])
            val = float(oloss.remfx_loss(m(x).squeeze(1), y))
            for k in keys:
                params[k].sub_(eps * dirs[k])
        return val

    eps = 1e-7  # the weights are O(1e-2): a larger step leaves the linear regime (central difference, fp64)
    fd = (loss_at(eps) - loss_at(-eps)) / (2 * eps)
    an = sum(float((r["param_grads"][k] * dirs[k]).sum()) for k in keys)
    assert abs(fd - an) < 1e-5 * max(1.0, abs(an)), (fd, an)


def test_oracles_match_reference_on_example_wav():
    """Real audio: the oracle restatements against the unchanged reference modules' outputs on example.wav."""
    from oracle import cnn14 as ocnn
    from oracle import hdemucs as ohd
    from tests.util import example_case

    g = golden("example_wav.npz")
    x, D = example_case(g)
    assert abs(float(x.pow(2).mean().sqrt()) - 0.1012) < 1e-3      # the file's level (RMS 0.101)
    sdu = weights.umx_state(0)
    assert abs(weights.checksum(sdu) - float(g["umx_wsum"])) < 1e-6 * abs(float(g["umx_wsum"]))
    assert relrms(oumx.sample(x, sdu)[0, 0, ::D], torch.from_numpy(g["umx_out"])) < 1e-5
    sdt = weights.tcn_state(0)
    out = otcn.sample(x[..., :int(g["tcn_T"])], sdt)
    assert out.shape[-1] == int(g["tcn_len"])
    assert relrms(out[0, 0, ::D], torch.from_numpy(g["tcn_out"])) < 1e-5
    sdc = weights.cnn14_state(0)
    lg = ocnn.logits(x, sdc)
    assert (lg - torch.from_numpy(g["logits"])).abs().max() < 1e-3
    assert torch.equal(ocnn.decisions(x, sdc), torch.from_numpy(g["decisions"]).long())
    assert relrms(ohd.sample(x, ohd.build(0))[0, 0, ::D], torch.from_numpy(g["hdemucs_out"])) < 1e-5


def _oracle_chain(x, y, member_seed0, use_all=False):
    from oracle import chain as ochain
    from oracle import cnn14 as ocnn
    from oracle.make_golden import CHAIN_ORDER

    sds = {e: weights.umx_state(member_seed0 + i) for i, e in enumerate(ochain.ALL_EFFECTS)}
    members = {e: (lambda sd: (lambda z: oumx.sample(z, sd)))(sd) for e, sd in sds.items()}
    csd = weights.cnn14_state(0)
    return ochain.forward(x, y, None, members, list(CHAIN_ORDER), classify=lambda z: torch.hstack(ocnn.forward(z, csd)), use_all=use_all)


def test_chain_oracle_matches_reference_golden():
    """oracle/chain.py against the UNCHANGED RemFXChainInference.forward (tests/golden/chain_forward.npz): detected labels,
    per-item cascade in cfg order, loss."""
    g = golden("chain_forward.npz")
    x, y = weights.synth_diverse(int(g["xseed"]), int(g["B"]), int(g["T"])), weights.synth_audio(int(g["yseed"]), int(g["B"]), int(g["T"]))
    loss, out, labels = _oracle_chain(x, y, int(g["member_seed0"]))
    assert torch.equal(labels, torch.from_numpy(g["labels"]))
    assert 0 < labels.sum() < labels.numel() and len({tuple(r.tolist()) for r in labels}) > 1   # the items take different paths
    assert relrms(out[:, 0, ::int(g["decim"])], torch.from_numpy(g["out"])) < 1e-5
    assert abs(float(loss) - float(g["loss"])) < 1e-5 * abs(float(g["loss"]))


@pytest.mark.skipif(not refshim.available(), reason="/root/reference not present (GPU box)")
def test_chain_oracle_matches_live_reference_use_all():
    """`use_all_effect_models=True` (remfx/models.py:65-69) live against the reference: every member applied to every item."""
    from oracle.make_golden import reference_chain

    x, y = weights.synth_diverse(90, 2, 32768), weights.synth_audio(91, 2, 32768)
    rloss, rout, _, _ = reference_chain(refshim.ref_modules(), x, y, 60, use_all=True)
    loss, out, _ = _oracle_chain(x, y, 60, use_all=True)
    assert relrms(out, rout) < 1e-5 and abs(float(loss) - float(rloss)) < 1e-5 * abs(float(rloss))


@pytest.mark.skipif(not refshim.available(), reason="/root/reference not present (GPU box)")
def test_umx_train_oracle_matches_reference():
    """oracle/umx_train.py against the UNCHANGED remfx.models.OpenUnmixModel in training mode (remfx/models.py:294-301): loss, output,
    every parameter gradient and the BatchNorm running statistics after one step (they move twice: the pass on spectrogram(x), then
    the separator pass).  The reference's LSTM dropout is set to 0 on the instance (no RNG stream can be shared); the masked form of
    the oracle is checked against itself below."""
    from oracle import umx_train as outr

    R = refshim.ref_modules()
    sd = weights.umx_state(5)
    m = R.models.OpenUnmixModel(n_fft=2048, hop_length=512, n_channels=1, alpha=0.3, sample_rate=48000)
    m.load_state_dict(sd, strict=True)
    m.train()
    m.model.lstm.dropout = 0.0
    x, t = weights.synth_audio(61, 2, 8192), weights.synth_audio(62, 2, 8192)
    loss, out = m((x, t))
    loss.backward()
    oloss_, oout, grads, stats = outr.train_grads((x, t), sd, None, None, dtype=torch.float64)
    assert abs(float(loss.detach()) - float(oloss_)) < 2e-5 * abs(float(oloss_))
    assert relrms(oout, out.detach()) < 2e-5
    for k, p in m.model.named_parameters():
        assert p.grad is not None, k
        if k == "input_mean":  # exactly zero: a constant added to every row of fc1's input is removed by bn1's batch mean
            assert float(p.grad.norm()) < 1e-5 * float(m.model.input_scale.grad.norm())
            assert float(grads[k].norm()) < 1e-9 * float(grads["input_scale"].norm())
            continue
        assert relrms(grads[k], p.grad) < 2e-3, (k, relrms(grads[k], p.grad))  # the reference runs in fp32, the oracle here in fp64
    for bn in ("bn1", "bn2", "bn3"):
        mod = getattr(m.model, bn)
        assert int(mod.num_batches_tracked) == 2
        assert relrms(stats[bn + ".running_mean"], mod.running_mean) < 1e-5, bn
        assert relrms(stats[bn + ".running_var"], mod.running_var) < 1e-5, bn
    # masks: an all-ones mask equals no mask; a real mask changes the loss
    ones = torch.ones(2, 2 * 17, 512)
    l1, _, _, _ = outr.train_grads((x, t), sd, ones, ones, dtype=torch.float64)
    assert abs(float(l1) - float(oloss_)) < 1e-9 * abs(float(oloss_))
    g = torch.Generator().manual_seed(0)
    msk = (torch.rand(2, 2 * 17, 512, generator=g) >= 0.4).float() / 0.6
    l2, _, _, _ = outr.train_grads((x, t), sd, msk, msk, dtype=torch.float64)
    assert abs(float(l2) - float(oloss_)) > 1e-6 * abs(float(oloss_))

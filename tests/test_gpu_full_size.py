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


def test_tcn_20_blocks_full_chunk_matches_reference_golden():
    """BASELINE config 1 at its real size: 20 blocks, 1 x 262144 (the TCN had only been compared up to 65536 samples).
    tests/golden/tcn_full_size.npz = the UNCHANGED reference TCNModel on example.wav (oracle/make_golden.py:tcn_full_size_golden),
    every 16th output sample."""
    from remfx_b200.models import TCNModel
    from tests.util import example_case, golden

    g = golden("tcn_full_size.npz")
    x, _ = example_case(golden("example_wav.npz"))
    sd = weights.tcn_state(0)
    assert abs(weights.checksum(sd) - float(g["tcn_wsum"])) < 1e-6 * abs(float(g["tcn_wsum"]))
    m = TCNModel(sample_rate=48000, num_bins=1025, ninputs=1, noutputs=1, nblocks=20, channel_growth=0, channel_width=256, kernel_size=7,
                 stack_size=10, dilation_growth=2, condition=False, latent_dim=2, norm_type="identity", causal=False, estimate_loudness=False)
    m.load_state_dict(sd, strict=True)
    out = m.cuda().eval().sample(x.cuda())
    assert out.shape == (1, 1, int(g["out_len"])) and out.shape[-1] == T - 12276
    err = relrms(out[0, 0, ::int(g["decim"])], torch.from_numpy(g["out"]))
    assert err < 1e-4, err


def test_hdemucs_batch_32_items_are_independent_of_their_batch_slot():
    """Hybrid Demucs at 32 x 262144 (config 3's batch): every normalisation is per item (TA:_hdemucs.py:554-563, GroupNorm), so a
    permuted batch gives the permuted outputs, and items 0 / 17 equal the same chunks run at B = 1, which
    test_gpu_hdemucs.py / test_gpu_zz_example_wav.py gate against torchaudio at 1e-4 -- stated here so that the B = 32 claim does not
    rest on an unstated argument.  One item is also compared with the torchaudio oracle directly."""
    from oracle import hdemucs as ohd
    from remfx_b200.models import DemucsModel

    ref = ohd.build(0)
    m = DemucsModel(sample_rate=48000, **ohd.KW)
    m.model.load_state_dict(ref.state_dict(), strict=True)
    m = m.cuda().eval()
    x = weights.synth_audio(9, B, T)
    xd = x.cuda()
    out = m.sample(xd)
    perm = torch.randperm(B, generator=torch.Generator().manual_seed(1)).cuda()
    out_p = m.sample(xd[perm].contiguous())
    # fp32 atomics in the GroupNorm statistics are order-dependent only within rounding
    assert relrms(out_p, out[perm]) < 2e-6
    for i in (0, 17):
        single = m.sample(xd[i:i + 1].contiguous())
        assert relrms(single, out[i:i + 1]) < 2e-6, i
    assert relrms(out[17:18].cpu(), ohd.sample(x[17:18], ref)) < 1e-4


def test_chain_16_full_chunks_with_hybrid_demucs_members():
    """BASELINE config 4 at its size: 16 x 262144 through classifier + cascade with the shipped architecture mix where an oracle
    exists -- Hybrid Demucs for distortion and compressor (cfg/exp/remfx_detect.yaml:63-68), Open-Unmix standing in for the three
    DCUNet members -- against the item-by-item oracle (oracle/chain.py over torchaudio HDemucs + the pinned Open-Unmix / Cnn14
    restatements).  Labels must be identical, outputs within 1e-4 (items 0..15 all compared)."""
    from oracle import chain as ochain
    from oracle import cnn14 as ocnn
    from oracle import hdemucs as ohd
    from remfx_b200.chain import RemFXChainInference
    from remfx_b200.classifier import Cnn14
    from remfx_b200.models import DemucsModel, OpenUnmixModel

    order = ["RandomPedalboardDistortion", "RandomPedalboardCompressor", "RandomPedalboardReverb", "RandomPedalboardChorus",
             "RandomPedalboardDelay"]
    Bc = 16
    members, omem = {}, {}
    for i, e in enumerate(ochain.ALL_EFFECTS):
        if e in ("RandomPedalboardDistortion", "RandomPedalboardCompressor"):
            ref = ohd.build(20 + i)
            mm = DemucsModel(sample_rate=48000, **ohd.KW)
            mm.model.load_state_dict(ref.state_dict(), strict=True)
            omem[e] = (lambda r: (lambda z: ohd.sample(z, r)))(ref)
        else:
            sd = weights.umx_state(50 + i)
            mm = OpenUnmixModel(sample_rate=48000)
            mm.load_state_dict(sd)
            omem[e] = (lambda s: (lambda z: oumx.sample(z, s)))(sd)
        members[e] = mm.cuda().eval()
    csd = weights.cnn14_state(0)
    clf = Cnn14(num_classes=5, sample_rate=48000, model_sample_rate=48000, n_fft=2048, hop_length=512, n_mels=128, specaugment=True)
    clf.load_state_dict(csd)
    clf = clf.cuda().eval()
    x, y = weights.synth_diverse(91, Bc, T), weights.synth_audio(92, Bc, T)
    chain = RemFXChainInference(members, 48000, 1025, order, classifier=clf)
    loss, out = chain((x.cuda(), y.cuda(), None, None), 0)
    torch.set_num_threads(max(1, torch.get_num_threads()))
    rloss, rout, rlabels = ochain.forward(x, y, None, omem, order, classify=lambda z: torch.hstack(ocnn.forward(z, csd)))
    assert torch.equal(chain.last_labels.cpu(), rlabels)
    assert 0 < float(rlabels.sum()) < Bc * 5  # the decisions actually vary over items / effects
    for i in range(Bc):
        assert relrms(out[i], rout[i]) < 1e-4, i
    assert abs(float(loss) - float(rloss)) < 1e-3 * abs(float(rloss))

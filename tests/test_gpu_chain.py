"""-m gpu: detect-then-remove cascade (remfx_b200.chain.RemFXChainInference) vs the item-by-item oracle."""
import pytest
import torch

from oracle import chain as ochain
from oracle import cnn14 as ocnn
from oracle import loss as oloss
from oracle import umx as oumx
from oracle import weights
from tests.util import relrms

pytestmark = pytest.mark.gpu

ORDER = ["RandomPedalboardDistortion", "RandomPedalboardCompressor", "RandomPedalboardReverb", "RandomPedalboardChorus",
         "RandomPedalboardDelay"]  # cfg/exp/remfx_detect.yaml:80-85


def _build(T):
    from remfx_b200.chain import RemFXChainInference
    from remfx_b200.classifier import Cnn14
    from remfx_b200.models import OpenUnmixModel

    sds = {e: weights.umx_state(50 + i) for i, e in enumerate(ochain.ALL_EFFECTS)}
    members = {}
    for e, sd in sds.items():
        m = OpenUnmixModel(sample_rate=48000)
        m.load_state_dict(sd)
        members[e] = m.cuda().eval()
    csd = weights.cnn14_state(0)
    clf = Cnn14(num_classes=5, sample_rate=48000, model_sample_rate=48000, n_fft=2048, hop_length=512, n_mels=128, specaugment=True)
    clf.load_state_dict(csd)
    clf = clf.cuda().eval()
    return sds, csd, members, clf, RemFXChainInference


@pytest.mark.parametrize("use_all", [False, True])
def test_chain_matches_oracle(use_all):
    T = 65536
    sds, csd, members, clf, Chain = _build(T)
    x, y = weights.synth_diverse(77, 4, T), weights.synth_audio(78, 4, T)
    chain = Chain(members, 48000, 1025, ORDER, classifier=clf, use_all_effect_models=use_all)
    loss, out = chain((x.cuda(), y.cuda(), None, None), 0)
    omem = {e: (lambda sd: (lambda z: oumx.sample(z, sd)))(sd) for e, sd in sds.items()}
    rloss, rout, rlabels = ochain.forward(x, y, None, omem, ORDER, classify=lambda z: torch.hstack(ocnn.forward(z, csd)), use_all=use_all)
    assert torch.equal(chain.last_labels.cpu(), rlabels)
    assert out.shape == rout.shape
    assert relrms(out, rout) < 1e-4
    assert abs(float(loss) - float(rloss)) < 1e-3 * abs(float(rloss))


def test_chain_with_given_labels_and_metrics():
    T = 32768
    sds, csd, members, clf, Chain = _build(T)
    x, y = weights.synth_audio(80, 3, T), weights.synth_audio(81, 3, T)
    labels = torch.tensor([[0., 0, 0, 1, 0], [1, 0, 0, 1, 1], [0, 0, 0, 0, 0]])
    chain = Chain(members, 48000, 1025, ORDER, classifier=None)
    loss, metrics = chain.test_step((x.cuda(), y.cuda(), None, labels), 0)
    omem = {e: (lambda sd: (lambda z: oumx.sample(z, sd)))(sd) for e, sd in sds.items()}
    rloss, rout, _ = ochain.forward(x, y, labels, omem, ORDER)
    assert abs(float(loss) - float(rloss)) < 1e-3 * abs(float(rloss))
    assert abs(float(metrics["test_SISDR"]) + float(oloss.sisdr_loss(rout, y))) < 1e-2
    assert abs(float(metrics["Input_SISDR"]) + float(oloss.sisdr_loss(x, y))) < 1e-3
    assert abs(float(metrics["Input_STFT"]) - float(oloss.mrstft(x, y))) < 1e-4 * float(oloss.mrstft(x, y))
    # item 2 has no effect detected: it must pass through untouched (models.py:96-104)
    out = chain.sample((x.cuda(), y.cuda(), None, labels))
    assert torch.equal(out[2].cpu(), x[2])


def test_whole_file_with_the_shipped_architecture_mix():
    """scripts/remfx_detect.py: ONE item of arbitrary length (a whole file, here 100003 samples -- off every grid the kernels tile by)
    through the classifier and the shipped member mix (cfg/exp/remfx_detect.yaml:63-78: Hybrid Demucs removes distortion and
    compressor, here Open-Unmix stands in for the three DCUNet members), every effect applied."""
    from oracle import hdemucs as ohd
    from remfx_b200.chain import RemFXChainInference
    from remfx_b200.classifier import Cnn14
    from remfx_b200.models import DemucsModel, OpenUnmixModel

    T = 100003
    refs, members, omem = {}, {}, {}
    for i, e in enumerate(ochain.ALL_EFFECTS):
        if e in ("RandomPedalboardDistortion", "RandomPedalboardCompressor"):
            ref = ohd.build(20 + i)
            m = DemucsModel(sample_rate=48000, **ohd.KW)
            m.model.load_state_dict(ref.state_dict(), strict=True)
            omem[e] = (lambda r: (lambda z: ohd.sample(z, r)))(ref)
        else:
            sd = weights.umx_state(60 + i)
            m = OpenUnmixModel(sample_rate=48000)
            m.load_state_dict(sd)
            omem[e] = (lambda sd: (lambda z: oumx.sample(z, sd)))(sd)
        members[e] = m.cuda().eval()
    csd = weights.cnn14_state(0)
    clf = Cnn14(num_classes=5, sample_rate=48000, model_sample_rate=48000, n_fft=2048, hop_length=512, n_mels=128, specaugment=True)
    clf.load_state_dict(csd)
    x = weights.synth_diverse(91, 1, T)
    chain = RemFXChainInference(members, 48000, 1025, ORDER, classifier=clf.cuda().eval(), use_all_effect_models=True)
    loss, out = chain((x.cuda(), x.cuda(), None, None), 0)
    rloss, rout, rlabels = ochain.forward(x, x, None, omem, ORDER, classify=lambda z: torch.hstack(ocnn.forward(z, csd)), use_all=True)
    assert torch.equal(chain.last_labels.cpu(), rlabels)
    assert out.shape == rout.shape == (1, 1, T)
    assert relrms(out, rout) < 2e-4   # five networks in sequence: the per-network 1e-4 gate compounds
    assert abs(float(loss) - float(rloss)) < 2e-3 * abs(float(rloss))


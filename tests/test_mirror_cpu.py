"""`not gpu`, live (skipped where /root/reference is absent): the host mirrors against the UNCHANGED reference classes --
state_dict keys, shapes and dtypes, checkpoint interchange in both directions, constructor-derived quantities, the crop helpers."""
import pytest
import torch

from oracle import refshim

pytestmark = pytest.mark.skipif(not refshim.available(), reason="/root/reference not present (GPU box)")

TCN_KW = dict(ninputs=1, noutputs=1, nblocks=20, channel_growth=0, channel_width=256, kernel_size=7, stack_size=10, dilation_growth=2,
              condition=False, latent_dim=2, norm_type="identity", causal=False, estimate_loudness=False)   # cfg/model/tcn.yaml:12-26


def _same_layout(mine, ref):
    a, b = mine.state_dict(), ref.state_dict()
    assert list(a.keys()) == list(b.keys())                     # same names in the same order
    for k in a:
        assert a[k].shape == b[k].shape and a[k].dtype == b[k].dtype, k
    mine.load_state_dict(b, strict=True)                        # a reference checkpoint loads into the drop-in ...
    ref.load_state_dict(mine.state_dict(), strict=True)         # ... and the drop-in's checkpoint loads into the reference
    for k, v in mine.state_dict().items():
        assert torch.equal(v, b[k]), k


def test_open_unmix_wrapper():
    from remfx_b200.models import OpenUnmixModel

    R = refshim.ref_modules()
    kw = dict(n_fft=2048, hop_length=512, n_channels=1, alpha=0.3, sample_rate=48000)               # cfg/model/umx.yaml:12-16
    _same_layout(OpenUnmixModel(**kw), R.models.OpenUnmixModel(**kw))


def test_tcn_wrapper():
    from remfx_b200.models import TCNModel

    R = refshim.ref_modules()
    mine, ref = TCNModel(sample_rate=48000, num_bins=1025, **TCN_KW), R.models.TCNModel(sample_rate=48000, num_bins=1025, **TCN_KW)
    _same_layout(mine, ref)
    assert mine.model.receptive_field == ref.model.receptive_field == 12277
    for k, p in mine.model.named_parameters():                                                        # same trainable set
        assert p.requires_grad == dict(ref.model.named_parameters())[k].requires_grad


def test_demucs_wrapper():
    from remfx_b200.models import DemucsModel

    R = refshim.ref_modules()
    kw = dict(sources=["mixture"], audio_channels=1, nfft=4096, channels=48)                           # cfg/model/demucs.yaml:12-16
    _same_layout(DemucsModel(sample_rate=48000, **kw), R.models.DemucsModel(sample_rate=48000, **kw))


def test_cnn14_classifier():
    from remfx_b200.classifier import Cnn14

    R = refshim.ref_modules()
    kw = dict(num_classes=5, n_fft=2048, hop_length=512, n_mels=128, sample_rate=48000, model_sample_rate=48000, specaugment=True)
    _same_layout(Cnn14(**kw), R.classifier.Cnn14(**kw))                                                # cfg/exp/remfx_detect.yaml:52-60


@pytest.mark.parametrize("L,l", [(100, 100), (100, 37), (262144, 249868), (9, 2)])
def test_crops(L, l):
    from remfx_b200.ops import causal_crop, center_crop

    R = refshim.ref_modules()
    x = torch.arange(2 * L, dtype=torch.float32).reshape(2, 1, L)
    assert torch.equal(center_crop(x, l), R.utils.center_crop(x, l))
    assert torch.equal(causal_crop(x, l), R.utils.causal_crop(x, l))   # incl. the reference's off-by-one (drops the last sample)


def test_label_order_is_the_reference_effect_list():
    from remfx_b200.chain import ALL_EFFECTS

    R = refshim.ref_modules()
    assert ALL_EFFECTS == [e.__name__ for e in R.models.ALL_EFFECTS]                                   # remfx/effects.py:699-705


def test_open_unmix_training_mode_host_logic():
    """Host side of OpenUnmixModel's training mode without a GPU: there is no CPU fallback (CPU tensors raise), the dropout masks have
    the shape / rate / scaling of nn.LSTM(dropout=0.4), and the running-statistics bookkeeping equals torch.nn.BatchNorm1d's."""
    from remfx_b200 import _lib
    from remfx_b200.models import OpenUnmixModel

    m = OpenUnmixModel(n_fft=2048, hop_length=512, n_channels=1, alpha=0.3, sample_rate=48000).train()
    with pytest.raises(_lib.RfxError):
        m((torch.zeros(1, 1, 8192), torch.zeros(1, 1, 8192)))
    torch.manual_seed(0)
    msk = m._dropout_masks(200, torch.device("cpu"))
    assert msk.shape == (2, 200, 512) and set(torch.unique(msk).tolist()) <= {0.0, float(torch.tensor(1.0) / torch.tensor(0.6))}
    assert abs(float((msk > 0).float().mean()) - 0.6) < 0.01
    m.model.lstm.dropout = 0.0
    assert m._dropout_masks(200, torch.device("cpu")) is None
    # running statistics: feed the same batch through torch's BatchNorm1d in training mode twice and through the mirror's update
    g = torch.Generator().manual_seed(1)
    rows = 37
    xs = [torch.randn(rows, n, generator=g) * 2 + 1 for n in (512, 512, 1025)]
    ref = [torch.nn.BatchNorm1d(n).train() for n in (512, 512, 1025)]
    stats = torch.cat([torch.cat([x.mean(0), x.var(0, unbiased=False)]) for x in xs])
    for _ in range(2):
        for bn, x in zip(ref, xs):
            bn(x)
        m._update_running_stats(stats, rows)
    for bn, mine in zip(ref, (m.model.bn1, m.model.bn2, m.model.bn3)):
        assert int(mine.num_batches_tracked) == int(bn.num_batches_tracked) == 2
        assert torch.allclose(mine.running_mean, bn.running_mean, atol=1e-6)
        assert torch.allclose(mine.running_var, bn.running_var, rtol=1e-5, atol=1e-6)

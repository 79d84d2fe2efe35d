"""-m gpu: TCN drop-in (remfx_b200.models.TCNModel) vs the reference golden and the oracle."""
import pytest
import torch

from oracle import tcn as otcn
from oracle import weights
from tests.util import golden, relrms

pytestmark = pytest.mark.gpu

TOL = 1e-4
KW = dict(ninputs=1, noutputs=1, nblocks=20, channel_growth=0, channel_width=256, kernel_size=7, stack_size=10,
          dilation_growth=2, condition=False, latent_dim=2, norm_type="identity", causal=False, estimate_loudness=False)


def _model(sd, **over):
    from remfx_b200.models import TCNModel

    kw = dict(KW)
    kw.update(over)
    m = TCNModel(sample_rate=48000, num_bins=1025, **kw)
    m.load_state_dict(sd, strict=True)
    return m.cuda().eval()


def test_forward_matches_reference_golden():
    g = golden("tcn_forward.npz")
    sd = weights.tcn_state(int(g["wseed"]))
    x = weights.synth_audio(int(g["xseed"]), 1, int(g["T"]))
    t = weights.synth_audio(int(g["tseed"]), 1, int(g["T"]))
    loss, out = _model(sd)((x.cuda(), t.cuda()))
    ref = torch.from_numpy(g["out"])
    assert out.shape == ref.shape == (1, 1, 4108)
    err = relrms(out, ref)
    assert err < TOL, err
    assert abs(float(loss) - float(g["loss"])) < 1e-3 * abs(float(g["loss"]))


@pytest.mark.parametrize("B,T,nblocks", [(2, 2000, 3), (3, 5000, 7), (1, 20000, 12)])
def test_sample_matches_oracle(B, T, nblocks):
    sd = weights.tcn_state(5, nblocks=nblocks)
    x = weights.synth_audio(200 + B, B, T)
    out = _model(sd, nblocks=nblocks).sample(x.cuda())
    ref = otcn.sample(x, sd)
    assert out.shape == ref.shape
    err = relrms(out, ref)
    assert err < TOL, err


def test_too_short_input_raises():
    sd = weights.tcn_state(5, nblocks=20)
    with pytest.raises(ValueError):
        _model(sd).sample(torch.zeros(1, 1, 4096, device="cuda"))

"""`not gpu`: the C-ABI library builds, loads and exports every symbol include/remfx_b200.h declares;
the Python drop-ins expose the reference's state_dict layout.  No compute calls (no GPU here)."""
import ctypes
import os
import re

import pytest
import torch

from oracle import weights

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def libpath():
    from remfx_b200 import build

    return build.build()


def test_header_symbols_are_exported(libpath):
    hdr = open(os.path.join(ROOT, "include", "remfx_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(rfx_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 10
    lib = ctypes.CDLL(libpath)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/remfx_b200.h but not exported"
    from remfx_b200 import _lib

    assert declared == set(_lib.EXPORTED_SYMBOLS), declared ^ set(_lib.EXPORTED_SYMBOLS)
    assert _lib.lib().rfx_abi_version() == 1


def test_umx_state_dict_layout_matches_reference():
    from remfx_b200.models import OpenUnmixModel

    m = OpenUnmixModel(n_fft=2048, hop_length=512, n_channels=1, alpha=0.3, sample_rate=48000)
    ref_keys = set(weights.umx_state(0).keys())  # pinned against the real reference in test_oracle_cpu / make_golden
    assert set(m.state_dict().keys()) == ref_keys
    m.load_state_dict(weights.umx_state(0), strict=True)
    assert m.model.fc1.weight is m.separator.target_models["other"].fc1.weight


def test_tcn_state_dict_layout_matches_reference():
    from remfx_b200.models import TCNModel

    m = TCNModel(sample_rate=48000, num_bins=1025, ninputs=1, noutputs=1, nblocks=20, channel_growth=0, channel_width=256,
                 kernel_size=7, stack_size=10, dilation_growth=2, condition=False, latent_dim=2, norm_type="identity", causal=False,
                 estimate_loudness=False)
    assert set(m.state_dict().keys()) == set(weights.tcn_state(0).keys())
    assert len(m.state_dict()) == 82
    assert m.model.receptive_field == 12277 and m.out_length(262144) == 249868


def test_cnn14_state_dict_layout_matches_reference():
    from remfx_b200.classifier import Cnn14

    m = Cnn14(num_classes=5, sample_rate=48000, model_sample_rate=48000, n_fft=2048, hop_length=512, n_mels=128, specaugment=True)
    ref = weights.cnn14_state(0, calibrate=False)
    assert set(m.state_dict().keys()) == set(ref.keys()) and len(ref) == 92
    m.load_state_dict(ref, strict=True)
    with pytest.raises(ValueError):
        Cnn14(num_classes=5, sample_rate=44100, model_sample_rate=48000)


def test_no_cpu_fallback():
    from remfx_b200 import _lib
    from remfx_b200.models import OpenUnmixModel

    m = OpenUnmixModel()
    with pytest.raises(_lib.RfxError):
        m.sample(torch.zeros(1, 1, 8192))


def test_tcn_training_path_refuses_cpu_tensors():
    """With trainable parameters and autograd on, TCNModel.forward takes the training path (rfx_tcn_forward_train /
    rfx_tcn_backward); like every product path it must refuse CPU tensors instead of computing somewhere else."""
    import pytest
    import torch

    from remfx_b200._lib import RfxError
    from remfx_b200.models import TCNModel

    m = TCNModel(sample_rate=48000, num_bins=1025, ninputs=1, noutputs=1, nblocks=2, channel_width=64, kernel_size=7, stack_size=10,
                 dilation_growth=2)
    x = torch.zeros(1, 1, 2000)
    assert any(p.requires_grad for p in m.parameters())
    with pytest.raises(RfxError):
        m((x, x))
    with torch.no_grad(), pytest.raises(RfxError):
        m((x, x))
    with pytest.raises(ValueError):
        m._sample_train(torch.zeros(1, 2, 2000))

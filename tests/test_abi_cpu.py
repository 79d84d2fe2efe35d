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


def test_tcn_training_abi_host_side_contract():
    """Host-only parts of the TCN training entry points (no kernel runs): workspace size formula, launch
    count, and loud refusal of an un-finalized handle."""
    import ctypes as C

    from remfx_b200 import _lib

    L = _lib.lib()
    cfg = _lib.TcnConfig(1, 1, 20, 256, 7, 10, 2, 0)
    h = C.c_void_p()
    _lib.check(L.rfx_tcn_create(C.byref(cfg), C.byref(h)), "rfx_tcn_create")
    try:
        B, T = 1, 262144
        L1 = T - 6
        plane = -(-(B * L1 * 256 * 2) // 256) * 256
        f32 = -(-(B * L1 * 256 * 4) // 256) * 256
        want = 20 * 2 * plane + 3 * f32 + 4 * plane + 8 * 256 * 256 * 4
        assert L.rfx_tcn_train_workspace_bytes(h, B, T) == want
        assert 6.2 < want / 2**30 < 6.3                                   # "6.4 GB per chunk" with parameters and loss buffers
        assert L.rfx_tcn_train_workspace_bytes(h, B, 12000) == 0          # shorter than the receptive field (12277)
        assert L.rfx_tcn_workspace_bytes(h, B, T) == 4 * plane            # inference: two ping-pong buffers
        assert L.rfx_tcn_backward_launches_per_call(h) == 2 + 5 * 19
        dummy = C.c_void_p(256)
        keys = (C.c_char_p * 1)(b"output.bias")
        ptrs = (C.c_void_p * 1)(256)
        rc = L.rfx_tcn_backward(h, dummy, dummy, dummy, B, T, keys, ptrs, 1, dummy, want, None)
        assert rc == 2 and b"finalize" in L.rfx_last_error()
        rc = L.rfx_tcn_forward_train(h, dummy, B, T, dummy, dummy, want, None)
        assert rc == 2 and b"finalize" in L.rfx_last_error()
        assert L.rfx_tcn_set_wgrad_impl(7) == 2 and L.rfx_tcn_set_wgrad_impl(0) == 0
    finally:
        L.rfx_tcn_destroy(h)


def test_precision_switch_and_hdemucs_training_entry_points_without_a_gpu():
    """Host-only behaviour of the round-2 entry points: the process-wide precision switch, and the Hybrid-Demucs training calls
    refusing an un-finalized handle / a workspace that does not belong to a recorded forward (no compute is launched)."""
    import ctypes as C

    import remfx_b200
    from remfx_b200 import _lib

    L = _lib.lib()
    assert remfx_b200.get_precision() == "fp32"
    remfx_b200.set_precision("bf16")
    assert L.rfx_get_matmul_precision() == 1 and remfx_b200.get_precision() == "bf16"
    remfx_b200.set_precision("fp32")
    assert L.rfx_get_matmul_precision() == 0
    assert L.rfx_set_matmul_precision(5) == 2 and b"mode" in L.rfx_last_error()
    with pytest.raises(ValueError):
        remfx_b200.set_precision("fp8")

    cfg = _lib.HDemucsConfig(1, 1, 48, 2, 4096, 6, 8, 4, 2, 1, 0, 4, 4, 2, 4, 4, 4, 0.2, 10.0)
    h = C.c_void_p()
    _lib.check(L.rfx_hdemucs_create(C.byref(cfg), C.byref(h)), "rfx_hdemucs_create")
    try:
        dummy = C.c_void_p(256)
        assert L.rfx_hdemucs_train_workspace_bytes(h, 2, 16384) == 0          # not finalized: no size
        rc = L.rfx_hdemucs_forward_train(h, dummy, 2, 16384, dummy, dummy, 1 << 30, None)
        assert rc == 2 and b"finalize" in L.rfx_last_error()
        keys = (C.c_char_p * 1)(b"freq_emb.embedding.weight")
        ptrs = (C.c_void_p * 1)(256)
        rc = L.rfx_hdemucs_backward(h, dummy, dummy, 2, 16384, keys, ptrs, 1, dummy, 1 << 30, None)
        assert rc == 2 and b"finalize" in L.rfx_last_error()
        dims = (C.c_int * 4)()
        assert L.rfx_hdemucs_grad_tap(h, b"freq_encoder.0", None, 0, dims, None) == 2   # no training forward has run
        assert L.rfx_hdemucs_inject_grad(h, b"freq_encoder.0", None) == 0
    finally:
        L.rfx_hdemucs_destroy(h)

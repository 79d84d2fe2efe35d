"""-m gpu: STFT / iSTFT kernels through the C ABI vs torch.stft/istft on the CPU and the golden KAT."""
import pytest
import torch

from oracle import stft as ostft
from oracle import weights
from tests.util import golden, relrms

pytestmark = pytest.mark.gpu


def _ops():
    from remfx_b200 import ops

    return ops


def test_stft_golden_kat():
    ops = _ops()
    g = golden("stft_kat.npz")
    x = weights.synth_audio(int(g["seed"]), int(g["B"]), int(g["T"]))
    Z = ops.stft(x.cuda(), 2048, 512, torch.hann_window(2048).cuda())  # (B,1,bins,F,2)
    assert Z.shape == (2, 1, 1025, 17, 2)
    assert relrms(Z[:, 0], torch.from_numpy(g["Z"])) < 2e-6
    y = ops.istft(Z, 2048, 512, torch.hann_window(2048).cuda(), length=8192)
    assert relrms(y[:, 0], torch.from_numpy(g["y"])) < 2e-6


@pytest.mark.parametrize("nfft", [512, 1024, 2048, 4096])
@pytest.mark.parametrize("hopdiv", [2, 4])
@pytest.mark.parametrize("T", [8192, 44100 // 2 * 2])
def test_roundtrip_and_torch(nfft, hopdiv, T):
    """umx/tests/test_transforms.py:42-51 (round trip RMSE < 1e-6) + elementwise vs torch.stft."""
    ops = _ops()
    hop = nfft // hopdiv
    T = T // hop * hop
    x = torch.rand(2, 1, T, generator=torch.Generator().manual_seed(nfft + T))
    win = torch.hann_window(nfft)
    Z = ops.stft(x.cuda(), nfft, hop, win.cuda())
    Zt = torch.view_as_real(torch.stft(x[:, 0], nfft, hop, window=win, return_complex=True))
    assert relrms(Z[:, 0], Zt) < 2e-6
    y = ops.istft(Z, nfft, hop, win.cuda(), length=T)
    assert float(torch.sqrt(((x - y.cpu()) ** 2).mean())) < 1e-6


def test_loss_resolutions_short_windows():
    ops = _ops()
    x = weights.synth_audio(5, 2, 16384)
    for n_fft, hop, win in [(1024, 120, 600), (2048, 240, 1200), (512, 50, 240)]:
        T = 16384 // hop * hop
        xs = x[:, 0, :T].contiguous()
        _, A = ops.stft_raw(xs.cuda(), n_fft, hop, torch.hann_window(win).cuda(), mode="mag_clamp", want_complex=False)
        Zt = torch.stft(xs, n_fft, hop, win, torch.hann_window(win), return_complex=True)
        ref = torch.sqrt(torch.clamp(Zt.real ** 2 + Zt.imag ** 2, min=1e-8)).transpose(1, 2)
        assert A.shape == ref.shape
        assert relrms(A, ref) < 2e-6


def test_spectrogram_matches_reference_formula():
    ops = _ops()
    x = weights.synth_audio(9, 2, 16384)
    S = ops.spectrogram(x.cuda(), torch.hann_window(2048).cuda(), 2048, 512, 0.3)
    ref = ostft.spectrogram(x, torch.hann_window(2048), 2048, 512, 0.3)
    assert S.shape == ref.shape
    assert relrms(S, ref) < 2e-6


def test_cpu_tensor_is_refused():
    ops = _ops()
    with pytest.raises(RuntimeError):
        ops.stft(torch.zeros(1, 1, 4096), 2048, 512)

"""Oracle: explicit restatement of torch.stft / torch.istft as the reference uses them.

TEST INFRASTRUCTURE (see oracle/__init__.py).

Follows
  * umx/openunmix/transforms.py:89-120  (TorchSTFT.forward: n_fft 2048, hop 512,
    hann periodic, center=True reflect, onesided, normalized=False)
  * umx/openunmix/transforms.py:164-181 (TorchISTFT.forward, length=T)
  * umx/openunmix/transforms.py:198-216 (ComplexNorm: abs)
  * remfx/utils.py:138-159              (spectrogram: (|STFT|+1e-8)^alpha)
  * remfx/utils.py:202-211              (center_crop / causal_crop)
The framing is written out (reflect pad, strided frames, window, rfft) rather
than calling torch.stft so that the kernels have an operation-level spec; the
`not gpu` tests check it against torch.stft / torch.istft bit-for-bit-ish
(SURVEY.md Appendix F).
"""
from __future__ import annotations

import torch


def hann_periodic(n: int, dtype=torch.float32) -> torch.Tensor:
    """torch.hann_window(n) (periodic): 0.5 - 0.5 cos(2 pi k / n) -- the call the reference makes
    (remfx/models.py:274, umx/openunmix/transforms.py:82)."""
    return torch.hann_window(n, periodic=True, dtype=dtype)


def padded_window(win_length: int, n_fft: int, dtype=torch.float32) -> torch.Tensor:
    """Hann(win_length) zero-padded (centred) to n_fft, as torch.stft does."""
    w = hann_periodic(win_length, dtype)
    if win_length == n_fft:
        return w
    left = (n_fft - win_length) // 2
    out = torch.zeros(n_fft, dtype=dtype)
    out[left : left + win_length] = w
    return out


def reflect_pad(x: torch.Tensor, pad: int) -> torch.Tensor:
    """x: (N, L) -> (N, L + 2 pad); reflection without repeating the edge sample."""
    left = x[:, 1 : pad + 1].flip(-1)
    right = x[:, -pad - 1 : -1].flip(-1)
    return torch.cat([left, x, right], dim=-1)


def stft(x: torch.Tensor, n_fft: int, hop: int, window: torch.Tensor, normalized: bool = False) -> torch.Tensor:
    """x: (N, L) real -> (N, n_fft/2+1, frames) complex64; center=True, reflect."""
    xp = reflect_pad(x, n_fft // 2)
    frames = xp.unfold(-1, n_fft, hop)  # (N, F, n_fft)
    spec = torch.fft.rfft(frames * window, dim=-1)  # (N, F, bins)
    if normalized:
        spec = spec * (n_fft ** -0.5)
    return spec.transpose(1, 2)


def istft(Z: torch.Tensor, n_fft: int, hop: int, window: torch.Tensor, length: int, normalized: bool = False) -> torch.Tensor:
    """Z: (N, bins, F) complex -> (N, length); center=True."""
    N, _, F = Z.shape
    if normalized:
        Z = Z * (n_fft ** 0.5)
    fr = torch.fft.irfft(Z.transpose(1, 2), n=n_fft, dim=-1) * window  # (N, F, n_fft)
    total = n_fft + hop * (F - 1)
    y = torch.zeros(N, total, dtype=fr.dtype)
    env = torch.zeros(total, dtype=fr.dtype)
    w2 = window * window
    for t in range(F):
        y[:, t * hop : t * hop + n_fft] += fr[:, t]
        env[t * hop : t * hop + n_fft] += w2
    start = n_fft // 2
    y = y[:, start : start + length]
    env = env[start : start + length]
    if y.shape[-1] < length:  # torch.istft zero-pads up to `length`
        padn = length - y.shape[-1]
        y = torch.nn.functional.pad(y, (0, padn))
        env = torch.nn.functional.pad(env, (0, padn), value=1.0)
    return y / env


def complex_norm(Z: torch.Tensor) -> torch.Tensor:
    return Z.abs()


def spectrogram(x: torch.Tensor, window: torch.Tensor, n_fft: int, hop: int, alpha: float) -> torch.Tensor:
    """remfx/utils.py:138-159; x: (B, C, T) -> (B, C, bins, frames)."""
    bs, chs, _ = x.shape
    X = stft(x.reshape(bs * chs, -1), n_fft, hop, window)
    X = X.reshape(bs, chs, X.shape[-2], X.shape[-1])
    return torch.pow(X.abs() + 1e-8, alpha)


def center_crop(x: torch.Tensor, length: int) -> torch.Tensor:
    start = (x.shape[-1] - length) // 2
    return x[..., start : start + length]


def causal_crop(x: torch.Tensor, length: int) -> torch.Tensor:
    """NB: drops the final sample (reference quirk)."""
    stop = x.shape[-1] - 1
    return x[..., stop - length : stop]

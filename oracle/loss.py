"""Oracle: auraloss MultiResolutionSTFTLoss / SISDRLoss restatement -- PARITY UNPINNED.

TEST INFRASTRUCTURE (see oracle/__init__.py).

`auraloss` is an un-vendored, un-pinned dependency of the reference
(setup.py:43; call sites remfx/models.py:7-8,35-44,171-176,289-292,312-315,
374-377) and is not installed in this image, so this file restates the
published auraloss>=0.4 algorithm with the defaults RemFx uses
(scale=None, w_sc=1, w_log_mag=1, w_lin_mag=0, w_phs=0):

  fft sizes (1024, 2048, 512), hops (120, 240, 50), win lengths (600, 1200, 240),
  hann windows; mag = sqrt(clamp(re^2 + im^2, 1e-8));
  SC  = ||Y - X||_F / ||Y||_F   per item, then mean over items
  LM  = mean |log X - log Y|
  loss = mean over the 3 resolutions of (SC + LM)

None of the reference's tests touches it: "parity unpinned" (DESIGN.md).
Sanity anchors: mrstft(a, a) = 0, mrstft(0.5a, a) = 0.5 + ln 2.
"""
from __future__ import annotations

import torch
from torch import nn

from oracle import stft as ostft

RESOLUTIONS = [(1024, 120, 600), (2048, 240, 1200), (512, 50, 240)]  # (n_fft, hop, win)


def stft_mag(x: torch.Tensor, n_fft: int, hop: int, win: int) -> torch.Tensor:
    """x: (N, T) -> (N, bins, frames) clamped magnitude."""
    w = ostft.padded_window(win, n_fft, x.dtype)
    X = ostft.stft(x, n_fft, hop, w)
    return torch.sqrt(torch.clamp(X.real ** 2 + X.imag ** 2, min=1e-8))


def mrstft(inp: torch.Tensor, tgt: torch.Tensor) -> torch.Tensor:
    x = inp.reshape(-1, inp.shape[-1])
    y = tgt.reshape(-1, tgt.shape[-1])
    total = 0.0
    for n_fft, hop, win in RESOLUTIONS:
        xm, ym = stft_mag(x, n_fft, hop, win), stft_mag(y, n_fft, hop, win)
        sc = (torch.linalg.norm(ym - xm, dim=(-2, -1)) / torch.linalg.norm(ym, dim=(-2, -1))).mean()
        lm = torch.nn.functional.l1_loss(torch.log(xm), torch.log(ym))
        total = total + sc + lm
    return total / len(RESOLUTIONS)


def sisdr_loss(inp: torch.Tensor, tgt: torch.Tensor, eps: float = 1e-8) -> torch.Tensor:
    """SISDRLoss(zero_mean=True, reduction='mean'): returns NEGATIVE SI-SDR in dB."""
    inp = inp - inp.mean(-1, keepdim=True)
    tgt = tgt - tgt.mean(-1, keepdim=True)
    alpha = (inp * tgt).sum(-1) / ((tgt ** 2).sum(-1) + eps)
    t = tgt * alpha.unsqueeze(-1)
    res = inp - t
    return -(10 * torch.log10((t ** 2).sum(-1) / ((res ** 2).sum(-1) + eps) + eps)).mean()


def remfx_loss(out: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """remfx/models.py:299,320,385: MRSTFT(out, target) + 100 * L1(out, target)."""
    return mrstft(out, target) + 100.0 * torch.nn.functional.l1_loss(out, target)


class MultiResolutionSTFTLoss(nn.Module):
    """Stand-in bound to `auraloss.freq.MultiResolutionSTFTLoss` by oracle/refshim.py."""

    def __init__(self, *a, **k):
        super().__init__()

    def forward(self, inp, tgt):
        return mrstft(inp, tgt)


class SISDRLoss(nn.Module):
    """Stand-in bound to `auraloss.time.SISDRLoss` by oracle/refshim.py."""

    def __init__(self, *a, **k):
        super().__init__()

    def forward(self, inp, tgt):
        return sisdr_loss(inp, tgt)

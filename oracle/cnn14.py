"""Oracle: Cnn14 effect classifier (eval mode), restated on torch-CPU.

TEST INFRASTRUCTURE (see oracle/__init__.py).

Follows
  * remfx/classifier.py:193-233  Cnn14.forward: MelSpectrogram -> per-item standardise -> 6 ConvBlocks
                                 -> mean over time -> max + mean over mel -> relu(fc1) -> 5 x sigmoid(Linear)
  * remfx/classifier.py:269-284  ConvBlock.forward: relu(bn1(conv3x3)) -> relu(bn2(conv3x3)) -> avg_pool2d
  * torchaudio.transforms.MelSpectrogram(sr, n_fft=2048, hop=512, n_mels=128): power STFT (hann, center,
    reflect, power=2) times the HTK filterbank `melspec.mel_scale.fb` (1025 x 128) from the state dict
  * remfx/models.py:61-64        decision = probability > 0.5  (== logit > 0)
`state` is the state_dict of `remfx.classifier.Cnn14` (conv_block{1..6}.{conv1,conv2,bn1,bn2}.*, fc1.*, heads.k.*).
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F

from oracle import stft as ostft

BN_EPS = 1e-5
POOLS = [(2, 2)] * 5 + [(1, 1)]


def melspec(x: torch.Tensor, state: Dict[str, torch.Tensor], n_fft: int = 2048, hop: int = 512) -> torch.Tensor:
    """x: (B, 1, T) -> (B, 1, n_mels, frames): fb^T |STFT|^2 (power, not log: remfx/classifier.py:200)."""
    B, C, T = x.shape
    Z = ostft.stft(x.reshape(B * C, T), n_fft, hop, state["melspec.spectrogram.window"])
    P = Z.real ** 2 + Z.imag ** 2  # (B, bins, frames)
    mel = torch.matmul(P.transpose(1, 2), state["melspec.mel_scale.fb"]).transpose(1, 2)
    return mel.reshape(B, C, mel.shape[-2], mel.shape[-1])


def _bn2d(x, state, p):
    rm, rv = state[p + ".running_mean"], state[p + ".running_var"]
    g, b = state[p + ".weight"], state[p + ".bias"]
    return (x - rm[None, :, None, None]) / torch.sqrt(rv[None, :, None, None] + BN_EPS) * g[None, :, None, None] + b[None, :, None, None]


def features(x: torch.Tensor, state) -> torch.Tensor:
    """(B, 1, T) -> (B, 2048) embedding after relu(fc1)."""
    h = melspec(x, state)
    h = (h - h.mean(dim=(2, 3), keepdim=True)) / h.std(dim=(2, 3), keepdim=True)  # unbiased std, no eps (classifier.py:207)
    for i in range(6):
        p = f"conv_block{i + 1}"
        h = torch.relu(_bn2d(F.conv2d(h, state[p + ".conv1.weight"], padding=1), state, p + ".bn1"))
        h = torch.relu(_bn2d(F.conv2d(h, state[p + ".conv2.weight"], padding=1), state, p + ".bn2"))
        h = F.avg_pool2d(h, kernel_size=POOLS[i])
    h = h.mean(dim=3)
    h = h.max(dim=2).values + h.mean(dim=2)
    return torch.relu(h @ state["fc1.weight"].t() + state["fc1.bias"])


def logits(x: torch.Tensor, state, num_classes: int = 5) -> torch.Tensor:
    e = features(x, state)
    return torch.cat([e @ state[f"heads.{k}.weight"].t() + state[f"heads.{k}.bias"] for k in range(num_classes)], dim=1)


def forward(x: torch.Tensor, state, num_classes: int = 5):
    """list of `num_classes` (B, 1) probabilities, as Cnn14.forward returns."""
    lg = logits(x, state, num_classes)
    return [torch.sigmoid(lg[:, k : k + 1]) for k in range(num_classes)]


def decisions(x: torch.Tensor, state, num_classes: int = 5) -> torch.Tensor:
    """remfx/models.py:63-64: where(hstack(probs) > 0.5, 1, 0) -> (B, num_classes) int."""
    return (torch.hstack(forward(x, state, num_classes)) > 0.5).to(torch.int64)

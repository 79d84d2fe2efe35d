"""Synthetic benchmark inputs (no datasets are reachable offline): clamp(0.1 N(0,1), -1, 1) mono chunks,
RMS 0.1 like the -20 LUFS-normalised RemFx dataset (remfx/datasets.py:237); seed from cfg/config.yaml:7."""
import torch


def synth_audio(seed: int, B: int, T: int) -> torch.Tensor:
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return (0.1 * torch.randn(B, 1, T, generator=g)).clamp_(-1.0, 1.0)

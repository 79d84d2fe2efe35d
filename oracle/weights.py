"""Seeded synthetic weights and inputs shared by the golden generator, tests and bench.

TEST INFRASTRUCTURE (see oracle/__init__.py).

No checkpoints are reachable (no network), so parity runs on seeded random weights in the
reference state_dict layouts.  They are *conditioned* (SURVEY.md section 8d): BatchNorm running
statistics are randomised so eval-BN is a non-trivial affine, and scales are chosen so
activations neither die nor saturate -- otherwise parity would not exercise the layers.
All draws come from a private CPU `torch.Generator`, so the same (seed, shape) gives the
same tensors here and on the GPU box (same torch build).
"""
from __future__ import annotations

from collections import OrderedDict

import torch


def _gen(seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return g


def synth_audio(seed: int, B: int, T: int) -> torch.Tensor:
    """clamp(0.1 N(0,1), -1, 1), shape (B, 1, T) -- RMS 0.1 like the -20 LUFS dataset."""
    g = _gen(seed)
    return (0.1 * torch.randn(B, 1, T, generator=g)).clamp_(-1.0, 1.0)


def _uniform(g, shape, bound):
    return (torch.rand(shape, generator=g) * 2 - 1) * bound


def _bn(sd, g, prefix, n, dims_tracked=True):
    sd[prefix + ".weight"] = 1.0 + 0.1 * torch.randn(n, generator=g)
    sd[prefix + ".bias"] = 0.1 * torch.randn(n, generator=g)
    sd[prefix + ".running_mean"] = 0.1 * torch.randn(n, generator=g)
    sd[prefix + ".running_var"] = 0.5 + torch.rand(n, generator=g)
    if dims_tracked:
        sd[prefix + ".num_batches_tracked"] = torch.tensor(0, dtype=torch.long)


def umx_state(seed: int = 0, n_fft: int = 2048, hidden: int = 512, layers: int = 3, sample_rate: float = 48000.0):
    """State dict with the key layout of `remfx.models.OpenUnmixModel` (96 tensors; the
    Open-Unmix weights appear under `model.*` and again under
    `separator.target_models.other.*` because the module is shared)."""
    g = _gen(seed)
    bins = n_fft // 2 + 1
    H = hidden // 2
    core = OrderedDict()
    core["input_mean"] = -0.5 * torch.rand(bins, generator=g)
    core["input_scale"] = 0.5 + torch.rand(bins, generator=g)
    core["output_scale"] = 0.5 + torch.rand(bins, generator=g)
    core["output_mean"] = 0.5 + 0.5 * torch.rand(bins, generator=g)
    core["fc1.weight"] = _uniform(g, (hidden, bins), bins ** -0.5)
    _bn(core, g, "bn1", hidden)
    k = H ** -0.5
    for l in range(layers):
        for suffix in ("", "_reverse"):
            core[f"lstm.weight_ih_l{l}{suffix}"] = _uniform(g, (4 * H, hidden), k)
            core[f"lstm.weight_hh_l{l}{suffix}"] = _uniform(g, (4 * H, H), k)
            core[f"lstm.bias_ih_l{l}{suffix}"] = _uniform(g, (4 * H,), k)
            core[f"lstm.bias_hh_l{l}{suffix}"] = _uniform(g, (4 * H,), k)
    core["fc2.weight"] = _uniform(g, (hidden, 2 * hidden), (2 * hidden) ** -0.5)
    _bn(core, g, "bn2", hidden)
    core["fc3.weight"] = _uniform(g, (bins, hidden), hidden ** -0.5)
    _bn(core, g, "bn3", bins)
    sd = OrderedDict()
    window = torch.hann_window(n_fft)
    sd["window"] = window
    for k_, v in core.items():
        sd["model." + k_] = v
    sd["separator.sample_rate"] = torch.as_tensor(sample_rate)
    sd["separator.stft.window"] = window.clone()
    sd["separator.istft.window"] = window.clone()
    for k_, v in core.items():
        sd["separator.target_models.other." + k_] = v
    return sd


def tcn_state(seed: int = 0, nblocks: int = 20, width: int = 256, kernel: int = 7, ninputs: int = 1, noutputs: int = 1):
    """State dict with the key layout of `remfx.models.TCNModel` (`model.process_blocks.N.*`, `model.output.*`).
    Scales keep the residual stream O(1) over 20 blocks (default init lets it drift)."""
    g = _gen(seed)
    sd = OrderedDict()
    cin = ninputs
    for n in range(nblocks):
        p = f"model.process_blocks.{n}"
        fan = cin * kernel
        sd[p + ".conv1.weight"] = _uniform(g, (width, cin, kernel), fan ** -0.5)
        sd[p + ".conv1.bias"] = _uniform(g, (width,), fan ** -0.5)
        sd[p + ".res.weight"] = _uniform(g, (width, cin, 1), 0.7 * (3.0 / cin) ** 0.5)
        sd[p + ".relu.weight"] = 0.25 + 0.1 * torch.randn(width, generator=g)
        cin = width
    sd["model.output.weight"] = _uniform(g, (noutputs, cin, 1), cin ** -0.5)
    sd["model.output.bias"] = _uniform(g, (noutputs,), cin ** -0.5)
    return sd


def checksum(sd) -> float:
    """Order-dependent fp64 checksum of a state dict (guards golden files against RNG drift)."""
    tot = 0.0
    for i, (k, v) in enumerate(sd.items()):
        tot += (i + 1) * float(v.double().sum()) + float(v.double().abs().sum())
    return tot

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


def synth_diverse(seed: int, B: int, T: int) -> torch.Tensor:
    """Spectrally diverse test signals (B, 1, T): white / red noise, tone stacks, chirps, AM bursts at random
    gains -- white noise alone looks identical to the classifier after its per-item standardisation."""
    import math

    g = _gen(seed)
    t = torch.arange(T, dtype=torch.float32) / 48000.0
    out = []
    for i in range(B):
        kind = int(torch.randint(0, 5, (1,), generator=g))
        if kind == 0:
            s = torch.randn(T, generator=g)
        elif kind == 1:
            s = torch.cumsum(torch.randn(T, generator=g), 0)
            s = s - torch.nn.functional.avg_pool1d(s[None, None], 2047, 1, 1023, count_include_pad=False)[0, 0]
        elif kind == 2:
            f = 60.0 * (1.0 + 60.0 * torch.rand(6, generator=g))
            s = sum(torch.sin(2 * math.pi * fi * t + float(torch.rand(1, generator=g)) * 6.28) / (k + 1) for k, fi in enumerate(f))
        elif kind == 3:
            f0, f1 = 50.0 + 500.0 * float(torch.rand(1, generator=g)), 2000.0 + 15000.0 * float(torch.rand(1, generator=g))
            s = torch.sin(2 * math.pi * (f0 * t + 0.5 * (f1 - f0) * t * t / t[-1]))
        else:
            env = (torch.sin(2 * math.pi * (1.0 + 8.0 * float(torch.rand(1, generator=g))) * t) > 0.3).float()
            s = env * torch.randn(T, generator=g)
        s = s / s.abs().max().clamp_min(1e-6) * (0.02 + 0.3 * float(torch.rand(1, generator=g)))
        out.append(s + 0.002 * torch.randn(T, generator=g))
    return torch.stack(out)[:, None, :].contiguous()


_CNN14_CACHE: dict = {}


def cnn14_state(seed: int = 0, num_classes: int = 5, n_fft: int = 2048, n_mels: int = 128, sample_rate: int = 48000,
                calibrate: bool = True):
    """Cached per process (the calibration runs the oracle on 8 full-length chunks); returns a fresh shallow copy."""
    key = (seed, num_classes, n_fft, n_mels, sample_rate, calibrate)
    if key not in _CNN14_CACHE:
        _CNN14_CACHE[key] = _cnn14_state(seed, num_classes, n_fft, n_mels, sample_rate, calibrate)
    return OrderedDict(_CNN14_CACHE[key])


def _cnn14_state(seed, num_classes, n_fft, n_mels, sample_rate, calibrate):
    """State dict with the key layout of `remfx.classifier.Cnn14` (92 tensors).  Conditioned so that features do
    not wash out (default init gives 0.495-0.505 for every input): He-scaled convs, randomised
    BN statistics, head weights/biases scaled so the five logits spread over roughly +-3 and vary per input."""
    import torchaudio

    g = _gen(seed)
    sd = OrderedDict()
    win = torch.hann_window(n_fft)
    sd["window"] = win
    sd["melspec.spectrogram.window"] = win.clone()
    sd["melspec.mel_scale.fb"] = torchaudio.functional.melscale_fbanks(n_fft // 2 + 1, 0.0, float(sample_rate // 2), n_mels, sample_rate,
                                                                        norm=None, mel_scale="htk")
    _bn(sd, g, "bn0", n_mels)
    chans = [1, 64, 128, 256, 512, 1024, 2048]
    for i in range(6):
        cin, cout = chans[i], chans[i + 1]
        p = f"conv_block{i + 1}"
        sd[p + ".conv1.weight"] = torch.randn(cout, cin, 3, 3, generator=g) * (2.0 / (cin * 9)) ** 0.5
        sd[p + ".conv2.weight"] = torch.randn(cout, cout, 3, 3, generator=g) * (2.0 / (cout * 9)) ** 0.5
        _bn(sd, g, p + ".bn1", cout)
        _bn(sd, g, p + ".bn2", cout)
    sd["fc1.weight"] = torch.randn(2048, 2048, generator=g) * (2.0 / 2048) ** 0.5
    sd["fc1.bias"] = 0.1 * torch.randn(2048, generator=g)
    for k in range(num_classes):
        sd[f"heads.{k}.weight"] = torch.randn(1, 2048, generator=g) * (4.0 / 2048 ** 0.5)
        sd[f"heads.{k}.bias"] = torch.randn(1, generator=g)
    if calibrate:
        # centre and scale the heads on a small fixed calibration set so decisions sit near the 0.5 threshold
        from oracle import cnn14 as _c

        with torch.no_grad():
            emb = _c.features(synth_diverse(seed + 7919, 8, 262144), sd)
            ebar = emb.mean(0, keepdim=True)
            for k in range(num_classes):
                w = sd[f"heads.{k}.weight"]
                # remove the common-mode component: a logit must not be the small difference of two large numbers,
                # otherwise ANY fp32 implementation (including the reference's own) decides by rounding noise
                w = w - (w @ ebar.t()) / (ebar @ ebar.t()) * ebar
                z = (emb @ w.t())[:, 0]
                scale = 2.0 / float(z.std().clamp_min(1e-6))
                sd[f"heads.{k}.weight"] = w * scale
                sd[f"heads.{k}.bias"] = (-(z * scale).median()).reshape(1) + 0.3 * torch.randn(1, generator=g)
    return sd


def checksum(sd) -> float:
    """Order-dependent fp64 checksum of a state dict (guards golden files against RNG drift)."""
    tot = 0.0
    for i, (k, v) in enumerate(sd.items()):
        tot += (i + 1) * float(v.double().sum()) + float(v.double().abs().sum())
    return tot

"""Oracle: Open-Unmix path of RemFx (eval mode), restated on torch-CPU.

TEST INFRASTRUCTURE (see oracle/__init__.py).

Follows
  * umx/openunmix/model.py:107-166   OpenUnmix.forward (fc1/bn1/tanh, 3xBiLSTM, skip,
                                      fc2/bn2/relu, fc3/bn3, output scale/mean, relu * mix)
  * umx/openunmix/model.py:242-319   Separator.forward (STFT -> |.| -> model -> wiener(niter=0) -> iSTFT)
  * umx/openunmix/filtering.py:442-459 wiener(softmask=False, niter=0): y = spec * (cos, sin)(angle(mix))
  * remfx/models.py:294-304          OpenUnmixModel.forward / sample
`state` is the reference state_dict of `remfx.models.OpenUnmixModel`
(keys `model.fc1.weight`, `model.lstm.weight_ih_l0[_reverse]`, `model.bn1.running_mean`, ...).

Two LSTM evaluators: `lstm_explicit` (gate-by-gate loop, the spec for the CUDA
recurrent kernel) and `lstm_fast` (torch's fused CPU LSTM, used
for the timed cpu_baseline so the baseline is not handicapped by a Python loop).
"""
from __future__ import annotations

from typing import Dict

import torch

from oracle import loss as oloss
from oracle import stft as ostft

BN_EPS = 1e-5


def _bn_eval(x, state, prefix):
    """BatchNorm1d eval: (x - running_mean) / sqrt(running_var + eps) * gamma + beta."""
    rm, rv = state[prefix + ".running_mean"], state[prefix + ".running_var"]
    g, b = state[prefix + ".weight"], state[prefix + ".bias"]
    return (x - rm) / torch.sqrt(rv + BN_EPS) * g + b


def lstm_explicit(x: torch.Tensor, state: Dict[str, torch.Tensor], prefix: str, layers: int = 3) -> torch.Tensor:
    """x: (T, B, I) -> (T, B, 2H). Gate order i, f, g, o; zero initial state."""
    inp = x
    for l in range(layers):
        outs = []
        for suffix in ("", "_reverse"):
            w_ih = state[f"{prefix}.weight_ih_l{l}{suffix}"]
            w_hh = state[f"{prefix}.weight_hh_l{l}{suffix}"]
            bias = state[f"{prefix}.bias_ih_l{l}{suffix}"] + state[f"{prefix}.bias_hh_l{l}{suffix}"]
            H = w_hh.shape[1]
            T, B, _ = inp.shape
            pre = inp @ w_ih.t() + bias  # (T, B, 4H)
            h = torch.zeros(B, H, dtype=x.dtype)
            c = torch.zeros(B, H, dtype=x.dtype)
            hs = [None] * T
            order = range(T) if suffix == "" else range(T - 1, -1, -1)
            for t in order:
                g = pre[t] + h @ w_hh.t()
                i_, f_, g_, o_ = g.split(H, dim=-1)
                c = torch.sigmoid(f_) * c + torch.sigmoid(i_) * torch.tanh(g_)
                h = torch.sigmoid(o_) * torch.tanh(c)
                hs[t] = h
            outs.append(torch.stack(hs, 0))
        inp = torch.cat(outs, -1)
    return inp


def lstm_fast(x: torch.Tensor, state: Dict[str, torch.Tensor], prefix: str, layers: int = 3) -> torch.Tensor:
    flat = []
    for l in range(layers):
        for suffix in ("", "_reverse"):
            for nm in ("weight_ih", "weight_hh", "bias_ih", "bias_hh"):
                flat.append(state[f"{prefix}.{nm}_l{l}{suffix}"])
    H = flat[1].shape[1]
    B = x.shape[1]
    hx = (torch.zeros(2 * layers, B, H, dtype=x.dtype), torch.zeros(2 * layers, B, H, dtype=x.dtype))
    out, _, _ = torch._VF.lstm(x, hx, flat, True, layers, 0.0, False, True, False)
    return out


def openunmix_forward(X: torch.Tensor, state: Dict[str, torch.Tensor], prefix: str = "model", fast_lstm: bool = True) -> torch.Tensor:
    """X: (B, C=1, bins, frames) magnitude -> same shape (model.py:107-166), eval mode."""
    x = X.permute(3, 0, 1, 2)
    F_, B, C, bins = x.shape
    mix = x.clone()
    hidden = state[prefix + ".fc1.weight"].shape[0]
    x = (x + state[prefix + ".input_mean"]) * state[prefix + ".input_scale"]
    x = x.reshape(-1, C * bins) @ state[prefix + ".fc1.weight"].t()
    x = _bn_eval(x, state, prefix + ".bn1")
    x = torch.tanh(x.reshape(F_, B, hidden))
    lstm = (lstm_fast if fast_lstm else lstm_explicit)(x, state, prefix + ".lstm")
    x = torch.cat([x, lstm], -1)
    x = x.reshape(-1, x.shape[-1]) @ state[prefix + ".fc2.weight"].t()
    x = torch.relu(_bn_eval(x, state, prefix + ".bn2"))
    x = x @ state[prefix + ".fc3.weight"].t()
    x = _bn_eval(x, state, prefix + ".bn3")
    x = x.reshape(F_, B, C, bins)
    x = x * state[prefix + ".output_scale"] + state[prefix + ".output_mean"]
    x = torch.relu(x) * mix
    return x.permute(1, 2, 3, 0)


def separator_forward(audio: torch.Tensor, state: Dict[str, torch.Tensor], n_fft: int = 2048, hop: int = 512,
                      fast_lstm: bool = True, wiener_trig: bool = True) -> torch.Tensor:
    """audio: (B, 1, T) -> (B, 1, 1, T).  `wiener_trig=True` follows filtering.py:442-451 literally
    (atan2 / cos / sin); False uses the algebraically identical mask * STFT the CUDA path implements."""
    B, C, T = audio.shape
    assert C == 1
    win = ostft.hann_periodic(n_fft, audio.dtype)
    Z = ostft.stft(audio.reshape(B, T), n_fft, hop, win)  # (B, bins, F) complex
    X = ostft.complex_norm(Z).unsqueeze(1)  # (B, 1, bins, F)
    spec = openunmix_forward(X, state, "model", fast_lstm)[:, 0]  # (B, bins, F)
    if wiener_trig:
        angle = torch.atan2(Z.imag, Z.real)
        Y = torch.complex(spec * torch.cos(angle), spec * torch.sin(angle))
    else:
        mag = Z.abs()
        mask = torch.where(mag > 0, spec / mag, torch.zeros_like(mag))
        Y = Z * mask
    y = ostft.istft(Y, n_fft, hop, win, length=T)
    return y.reshape(B, 1, 1, T)


def sample(x: torch.Tensor, state, **kw) -> torch.Tensor:
    """remfx/models.py:303-304: separator(x).squeeze(1) -> (B, 1, T)."""
    return separator_forward(x, state, **kw).squeeze(1)


def forward(batch, state, **kw):
    """remfx/models.py:294-301 in eval mode (the dead `Y = self.model(X)` pass has no
    observable effect in eval): returns (loss, sep_out)."""
    x, target = batch
    out = sample(x, state, **kw)
    return oloss.remfx_loss(out, target), out

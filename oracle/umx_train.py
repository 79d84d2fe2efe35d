"""Oracle: the Open-Unmix TRAINING step of RemFx, restated on torch-CPU under autograd.

TEST INFRASTRUCTURE (see oracle/__init__.py).

Follows
  * remfx/models.py:294-301        OpenUnmixModel.forward in training mode: `X = spectrogram(x); Y = self.model(X)` (a pass whose output
                                    is discarded but which moves the BatchNorm running statistics), then `self.separator(x)`, then
                                    MRSTFT + 100 L1
  * remfx/utils.py:138-159         spectrogram = (|STFT| + 1e-8) ** alpha
  * umx/openunmix/model.py:107-166 OpenUnmix.forward with BatchNorm1d batch statistics (training) and nn.LSTM(dropout=0.4)
                                    (model.py:62-69: dropout on the output of every layer but the last)
  * umx/openunmix/model.py:242-319 Separator.forward: the network sees `X.detach().clone()`; wiener(niter=0) keeps the mixture phase
  * torch.nn.BatchNorm1d           running = 0.9 running + 0.1 batch (variance unbiased), num_batches_tracked += 1

Dropout masks are explicit inputs ((layers - 1, frames * batch, hidden), entries 0 or 1 / (1 - p), frame-major rows m = b * F + t as the
CUDA path stores them) because no two RNG streams agree; `None` = no dropout.  Pinned against the unchanged reference class in training
mode with its LSTM dropout set to 0 (tests/test_oracle_cpu.py::test_umx_train_oracle_matches_reference).
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from oracle import loss as oloss
from oracle import stft as ostft

BN_EPS = 1e-5
BN_MOMENTUM = 0.1


def _bn_train(x: torch.Tensor, state: Dict[str, torch.Tensor], prefix: str, new_stats: Dict[str, torch.Tensor]) -> torch.Tensor:
    """BatchNorm1d in training mode on (rows, features); records the running-statistics update in `new_stats`."""
    mean = x.mean(0)
    var_b = x.var(0, unbiased=False)
    n = x.shape[0]
    with torch.no_grad():
        rm = new_stats.get(prefix + ".running_mean", state[prefix + ".running_mean"].to(x.dtype))
        rv = new_stats.get(prefix + ".running_var", state[prefix + ".running_var"].to(x.dtype))
        new_stats[prefix + ".running_mean"] = (1 - BN_MOMENTUM) * rm + BN_MOMENTUM * mean.detach()
        new_stats[prefix + ".running_var"] = (1 - BN_MOMENTUM) * rv + BN_MOMENTUM * var_b.detach() * n / max(n - 1, 1)
    return (x - mean) / torch.sqrt(var_b + BN_EPS) * state[prefix + ".weight"] + state[prefix + ".bias"]


def _lstm_layer(x: torch.Tensor, state, prefix: str, l: int) -> torch.Tensor:
    """One bidirectional layer through torch's fused CPU LSTM (differentiable); x: (T, B, I) -> (T, B, 2H)."""
    flat = []
    for suffix in ("", "_reverse"):
        for nm in ("weight_ih", "weight_hh", "bias_ih", "bias_hh"):
            flat.append(state[f"{prefix}.{nm}_l{l}{suffix}"])
    H = flat[1].shape[1]
    B = x.shape[1]
    hx = (torch.zeros(2, B, H, dtype=x.dtype), torch.zeros(2, B, H, dtype=x.dtype))
    out, _, _ = torch._VF.lstm(x, hx, flat, True, 1, 0.0, False, True, False)
    return out


def network_train(X: torch.Tensor, state, masks: Optional[torch.Tensor], new_stats, prefix: str = "model", layers: int = 3) -> torch.Tensor:
    """X: (B, 1, bins, frames) -> same shape, training mode.  masks: (layers - 1, B * F, hidden) in b-major rows, or None."""
    x = X.permute(3, 0, 1, 2)
    F_, B, C, bins = x.shape
    mix = x.detach().clone()
    hidden = state[prefix + ".fc1.weight"].shape[0]
    x = (x + state[prefix + ".input_mean"]) * state[prefix + ".input_scale"]
    x = x.reshape(-1, C * bins) @ state[prefix + ".fc1.weight"].t()
    x = _bn_train(x, state, prefix + ".bn1", new_stats)
    x = torch.tanh(x.reshape(F_, B, hidden))
    inp = x
    for l in range(layers):
        inp = _lstm_layer(inp, state, prefix + ".lstm", l)
        if masks is not None and l + 1 < layers:
            m = masks[l].reshape(B, F_, hidden).permute(1, 0, 2).to(inp.dtype)  # rows b * F + t -> (t, b)
            inp = inp * m
    x = torch.cat([x, inp], -1)
    x = x.reshape(-1, x.shape[-1]) @ state[prefix + ".fc2.weight"].t()
    x = torch.relu(_bn_train(x, state, prefix + ".bn2", new_stats))
    x = x @ state[prefix + ".fc3.weight"].t()
    x = _bn_train(x, state, prefix + ".bn3", new_stats)
    x = x.reshape(F_, B, C, bins)
    x = x * state[prefix + ".output_scale"] + state[prefix + ".output_mean"]
    x = torch.relu(x) * mix
    return x.permute(1, 2, 3, 0)


def train_forward(batch, state, masks_dead: Optional[torch.Tensor] = None, masks_real: Optional[torch.Tensor] = None, n_fft: int = 2048,
                  hop: int = 512, alpha: float = 0.3, dead_pass: bool = True):
    """(x, target) -> (loss, sep_out, new_stats) with `state` holding leaf tensors that may require grad."""
    x, target = batch
    B, C, T = x.shape
    dtype = state["model.fc1.weight"].dtype
    x = x.to(dtype)
    target = target.to(dtype)
    win = ostft.hann_periodic(n_fft, dtype)
    new_stats: Dict[str, torch.Tensor] = {}
    Z = ostft.stft(x.reshape(B, T), n_fft, hop, win)  # (B, bins, F) complex
    if dead_pass:
        with torch.no_grad():
            Xp = torch.pow(Z.abs() + 1e-8, alpha).unsqueeze(1)
            network_train(Xp, state, masks_dead, new_stats)
    mag = Z.abs().unsqueeze(1)
    spec = network_train(mag.detach().clone(), state, masks_real, new_stats)[:, 0]  # (B, bins, F)
    angle = torch.atan2(Z.imag, Z.real)
    Y = torch.complex(spec * torch.cos(angle), spec * torch.sin(angle))
    out = ostft.istft(Y, n_fft, hop, win, length=T).reshape(B, 1, T)
    loss = oloss.remfx_loss(out, target)
    return loss, out, new_stats


def train_grads(batch, sd, masks_dead=None, masks_real=None, dtype=torch.float64, cotangent=None, **kw):
    """Loss, output, parameter gradients (keys without the `model.` prefix) and the updated running statistics of one training-mode
    forward + backward, evaluated in `dtype`.  `cotangent` (shaped like the output): differentiate <out, cotangent> instead of the
    loss -- a linear objective separates the network's backward from the conditioning of the loss gradient."""
    state = {}
    for k, v in sd.items():
        if not v.is_floating_point():
            continue
        t = v.detach().to(dtype).clone()
        if k.startswith("model.") and "running_" not in k:
            t.requires_grad_(True)
        state[k] = t
    loss, out, new_stats = train_forward(batch, state, masks_dead, masks_real, **kw)
    if cotangent is not None:
        (out * cotangent.to(dtype)).sum().backward()
    else:
        loss.backward()
    grads = {k[len("model."):]: v.grad.detach() for k, v in state.items() if v.requires_grad and v.grad is not None}
    return loss.detach(), out.detach(), grads, {k[len("model."):]: v for k, v in new_stats.items()}

"""Oracle: TCN path of RemFx, restated on torch-CPU.

TEST INFRASTRUCTURE (see oracle/__init__.py).

Follows
  * remfx/tcn.py:48-59    TCNBlock.forward: PReLU(conv1d(x; k, dilation d, no pad)) + center_crop(res1x1(x))
  * remfx/tcn.py:105-118  dilation = growth ** (n % stack_size); Cin = ninputs for n = 0
  * remfx/tcn.py:126-130  TCN.forward: blocks, then tanh(conv1d 1x1 + bias)
  * remfx/models.py:379-390 TCNModel.forward / sample (causal_crop of the target, loss)
`state` is the state_dict of `remfx.models.TCNModel` (`model.process_blocks.{n}.conv1.weight`, ...).

`block_fused` is the single-formula form the CUDA kernel implements: the residual's
centre crop offset is 3d (k=7), i.e. the residual is an 8th tap on the centre sample.
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F

from oracle import loss as oloss
from oracle import stft as ostft


def n_blocks(state: Dict[str, torch.Tensor], prefix: str = "model") -> int:
    n = 0
    while f"{prefix}.process_blocks.{n}.conv1.weight" in state:
        n += 1
    return n


def dilation_of(n: int, growth: int = 2, stack: int = 10) -> int:
    return growth ** (n % stack)


def block(x, state, prefix, n, growth=2, stack=10):
    p = f"{prefix}.process_blocks.{n}"
    d = dilation_of(n, growth, stack)
    y = F.conv1d(x, state[p + ".conv1.weight"], state[p + ".conv1.bias"], dilation=d)
    y = F.prelu(y, state[p + ".relu.weight"])
    r = F.conv1d(x, state[p + ".res.weight"])
    return y + ostft.center_crop(r, y.shape[-1])


def block_fused(x, state, prefix, n, growth=2, stack=10):
    """Same value as `block`, written as 7 dilated taps + an 8th centre tap (kernel spec)."""
    p = f"{prefix}.process_blocks.{n}"
    d = dilation_of(n, growth, stack)
    w, b = state[p + ".conv1.weight"], state[p + ".conv1.bias"]
    k = w.shape[-1]
    Lout = x.shape[-1] - (k - 1) * d
    acc = b.view(1, -1, 1).expand(x.shape[0], -1, Lout).clone()
    for j in range(k):
        acc = acc + torch.einsum("oc,bcl->bol", w[:, :, j], x[:, :, j * d : j * d + Lout])
    a = state[p + ".relu.weight"].view(1, -1, 1)
    y = torch.where(acc >= 0, acc, a * acc)
    off = ((k - 1) * d) // 2
    return y + torch.einsum("oc,bcl->bol", state[p + ".res.weight"][:, :, 0], x[:, :, off : off + Lout])


def tcn_forward(x: torch.Tensor, state, prefix: str = "model", growth=2, stack=10, fused=False) -> torch.Tensor:
    fn = block_fused if fused else block
    for n in range(n_blocks(state, prefix)):
        x = fn(x, state, prefix, n, growth, stack)
    return torch.tanh(F.conv1d(x, state[prefix + ".output.weight"], state[prefix + ".output.bias"]))


def sample(x, state, **kw):
    return tcn_forward(x, state, **kw)


def forward(batch, state, **kw):
    x, target = batch
    out = tcn_forward(x, state, **kw)
    if out.shape[-1] < target.shape[-1]:
        target = ostft.causal_crop(target, out.shape[-1])
    return oloss.remfx_loss(out, target), out

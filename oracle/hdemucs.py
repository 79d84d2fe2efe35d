"""Oracle: Hybrid Demucs as RemFx uses it -- PARITY UNPINNED UPSTREAM.

TEST INFRASTRUCTURE (see oracle/__init__.py).

RemFx's "Demucs" is `torchaudio.models.HDemucs` (remfx/models.py:6,308-324; cfg/model/demucs.yaml:11-16:
sources=["mixture"], audio_channels=1, nfft=4096, channels=48).  The source is a third-party dependency that the
reference does not vendor and does not pin (`torchaudio>=0.13.0`, setup.py:33), and none of the reference's tests
covers it, so there is nothing upstream to pin against: the oracle IS the torchaudio build of this image
(2.11.0), executed on the CPU.  Intermediate activations for layer-by-layer parity come from forward hooks.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict

import torch

KW = dict(sources=["mixture"], audio_channels=1, nfft=4096, channels=48)


def build(seed: int = 0, layerscale: float = 0.1, **over):
    """Seeded HDemucs in eval mode.  LayerScale starts at 1e-4, which would make every DConv branch (and with it
    the LSTMs and the local attention) numerically invisible; it is raised to `layerscale` so parity exercises them."""
    from torchaudio.models import HDemucs

    kw = dict(KW)
    kw.update(over)
    torch.manual_seed(seed)
    m = HDemucs(**kw)
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for name, p in m.named_parameters():
            if name.endswith(".scale"):  # _LayerScale
                p.copy_(layerscale * (0.5 + torch.rand(p.shape, generator=g)))
            elif "norm" in name.split(".")[-2] or name.split(".")[-2].isdigit() and p.dim() == 1 and name.endswith("weight"):
                pass
        for name, mod in m.named_modules():
            if isinstance(mod, torch.nn.GroupNorm):
                mod.weight.copy_(1.0 + 0.1 * torch.randn(mod.weight.shape, generator=g))
                mod.bias.copy_(0.1 * torch.randn(mod.bias.shape, generator=g))
    return m.eval()


def state(seed: int = 0, **kw) -> "OrderedDict[str, torch.Tensor]":
    return OrderedDict((k, v.clone()) for k, v in build(seed, **kw).state_dict().items())


def sample(x: torch.Tensor, m) -> torch.Tensor:
    """DemucsModel.sample (remfx/models.py:323-324): self.model(x).squeeze(1) -> (B, 1, T)."""
    with torch.no_grad():
        return m(x).squeeze(1)


def taps(x: torch.Tensor, m, names) -> Dict[str, torch.Tensor]:
    """Outputs of the named sub-modules during one forward (for layer-by-layer parity tests)."""
    out, hooks = {}, []
    mods = dict(m.named_modules())
    for n in names:
        hooks.append(mods[n].register_forward_hook(lambda mod, inp, o, n=n: out.__setitem__(n, (o[0] if isinstance(o, tuple) else o).detach().clone())))
    with torch.no_grad():
        out["__output__"] = m(x).squeeze(1)
    for h in hooks:
        h.remove()
    return out


def grad_taps(x: torch.Tensor, target: torch.Tensor, m, names=(), objective=None):
    """Backward oracle for the Hybrid-Demucs training step (remfx/models.py:317-321 under `loss.backward()`): runs the module
    under torch autograd and returns
      loss            the training loss MRSTFT + 100 L1 (oracle/loss.py) of the (B, 1, T) output, or `objective(out)` when given,
      output          (B, 1, T),
      param_grads     {state_dict key: gradient},
      act_grads       {sub-module name: dLoss / d(that sub-module's output)}  for layer-by-layer parity of a backward pass
                      (tensor hooks on the forward outputs of the named sub-modules; tuple outputs -> first element).
    The module is left with zeroed gradients."""
    from oracle import loss as oloss

    act_grads, hooks = {}, []
    mods = dict(m.named_modules())

    def fwd_hook(name):
        def hook(mod, inp, o):
            t = o[0] if isinstance(o, tuple) else o
            if t.requires_grad:
                t.register_hook(lambda g, name=name: act_grads.__setitem__(name, g.detach().clone()))
        return hook

    for n in names:
        hooks.append(mods[n].register_forward_hook(fwd_hook(n)))
    m.zero_grad(set_to_none=True)
    out = m(x).squeeze(1)
    loss = objective(out) if objective is not None else oloss.remfx_loss(out, target)
    loss.backward()
    for h in hooks:
        h.remove()
    param_grads = {k: p.grad.detach().clone() for k, p in m.named_parameters() if p.grad is not None}
    m.zero_grad(set_to_none=True)
    return dict(loss=loss.detach(), output=out.detach(), param_grads=param_grads, act_grads=act_grads)

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

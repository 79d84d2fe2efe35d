"""Oracle: RemFXChainInference.forward (remfx/models.py:52-108) restated item by item on torch-CPU.

TEST INFRASTRUCTURE (see oracle/__init__.py).  `members[effect]` is a callable `sample(x)` built from the other
oracles; `classify(x)` returns the (B, 5) probabilities (or None to use the given labels)."""
from __future__ import annotations

import torch

from oracle import loss as oloss

ALL_EFFECTS = ["RandomPedalboardReverb", "RandomPedalboardChorus", "RandomPedalboardDelay", "RandomPedalboardDistortion",
               "RandomPedalboardCompressor"]  # remfx/effects.py:699-705


def forward(x, y, labels, members, effects_order, classify=None, use_all=False):
    if classify is not None:
        labels = torch.where(classify(x) > 0.5, 1.0, 0.0)  # models.py:61-64
    out = []
    for i in range(x.shape[0]):
        elem = x[i : i + 1]
        names = [ALL_EFFECTS[j] for j in range(len(ALL_EFFECTS)) if use_all or labels[i, j] == 1.0]
        for effect in [e for e in effects_order if e in names]:  # models.py:96-103
            elem = members[effect](elem)
        out.append(elem[0])
    out = torch.stack(out)
    return oloss.remfx_loss(out, y), out, labels

"""Drop-in for `remfx.models.RemFXChainInference` (remfx/models.py:22-149): detect-then-remove cascade.

Same constructor arguments and `forward(batch, batch_idx, order=None, verbose=False) -> (loss, output)` /
`sample(batch)` as the reference.  The reference walks the batch item by item and applies the effect-specific
models one after another at batch size 1 (models.py:93-104); items are independent on this path, so here the
items that need effect E are gathered into ONE real batch per effect model, in `effects_order` -- the same
result with B-fold fewer launches.  Classifier decisions use the reference's hard-coded
0.5 threshold (models.py:61-64) and its label order (remfx/effects.py:699-705).
"""
from __future__ import annotations

import random
from typing import Dict, List, Optional, Sequence

import torch
from torch import Tensor, nn

from .losses import mrstft_loss, remfx_loss_with_terms, sisdr_loss
from .ops import causal_crop

# remfx/effects.py:699-705 (Pedalboard_Effects): label index -> effect class name
ALL_EFFECTS: List[str] = [
    "RandomPedalboardReverb",
    "RandomPedalboardChorus",
    "RandomPedalboardDelay",
    "RandomPedalboardDistortion",
    "RandomPedalboardCompressor",
]


def _network_of(member):
    """The reference stores Lightning `RemFX` modules and calls `.model.sample` (models.py:103); accept bare networks too."""
    inner = getattr(member, "model", None)
    if inner is not None and hasattr(inner, "sample"):
        return inner
    return member


class RemFXChainInference(nn.Module):
    def __init__(self, models: Dict[str, nn.Module], sample_rate, num_bins, effect_order: Sequence[str], classifier=None,
                 shuffle_effect_order: bool = False, use_all_effect_models: bool = False):
        super().__init__()
        self.model = models  # a plain dict, as in the reference (members are not registered sub-modules)
        self.sample_rate = sample_rate
        self.num_bins = num_bins
        self.effect_order = list(effect_order)
        self.classifier = classifier
        self.shuffle_effect_order = shuffle_effect_order
        self.use_all_effect_models = use_all_effect_models
        self.output_str = "IN_SISDR,OUT_SISDR,IN_STFT,OUT_STFT\n"
        self.last_labels: Optional[Tensor] = None

    def detect(self, x: Tensor) -> Tensor:
        """(B, 1, T) -> (B, 5) float 0/1 labels: where(hstack(classifier(x)) > 0.5, 1, 0) (models.py:61-64)."""
        with torch.no_grad():
            labels = torch.hstack(list(self.classifier(x)))
        return torch.where(labels > 0.5, 1.0, 0.0)

    def forward(self, batch, batch_idx: int = 0, order: Optional[Sequence[str]] = None, verbose: bool = False):
        x, y, _, rem_fx_labels = batch
        effects_order = list(order) if order else self.effect_order
        if self.classifier:
            rem_fx_labels = self.detect(x)
        rem_fx_labels = torch.as_tensor(rem_fx_labels)
        self.last_labels = rem_fx_labels
        if self.use_all_effect_models:
            present = torch.ones(x.shape[0], len(ALL_EFFECTS), dtype=torch.bool)
        else:
            present = (rem_fx_labels == 1.0).cpu()
        if verbose:
            print("Detected effects:", [ALL_EFFECTS[i] for i in range(len(ALL_EFFECTS)) if present[0, i]])
            print("Removing effects...")
        output = x
        cloned = False
        with torch.no_grad():
            for effect in effects_order:
                if effect not in ALL_EFFECTS:
                    continue
                idx = torch.nonzero(present[:, ALL_EFFECTS.index(effect)]).flatten()
                if idx.numel() == 0:
                    continue
                net = _network_of(self.model[effect])
                if idx.numel() == x.shape[0]:
                    output = net.sample(output.contiguous())
                    cloned = True
                else:
                    if not cloned:
                        output = output.clone()
                        cloned = True
                    sel = idx.to(output.device)
                    output[sel] = net.sample(output[sel].contiguous())
        loss, self.last_loss_terms = remfx_loss_with_terms(output, y)
        return loss, output

    def sample(self, batch):
        return self.forward(batch, 0)[1]

    def test_step(self, batch, batch_idx: int = 0):
        """models.py:110-145 without the Lightning logger: returns (loss, metrics dict) with the reference's metric names."""
        x, y, _, _ = batch
        if self.shuffle_effect_order:
            random.shuffle(self.effect_order)
        loss, output = self.forward(batch, batch_idx, order=self.effect_order)
        cropped = output.shape[-1] < y.shape[-1]
        if cropped:
            y = causal_crop(y, output.shape[-1])
        terms = None if cropped else self.__dict__.get("last_loss_terms")  # MR-STFT of (output, y): a by-product of the loss kernels
        with torch.no_grad():
            metrics = {
                "test_loss": loss,
                "test_SISDR": -sisdr_loss(output, y),
                "Input_SISDR": -sisdr_loss(x, y),
                "test_STFT": terms[1] if terms is not None else mrstft_loss(output, y),
                "Input_STFT": mrstft_loss(x, y),
            }
        return loss, metrics

"""Rows L4 / L5: the reference's training harness step around a network wrapper, without Lightning.

Mirror of `remfx.models.RemFX` (remfx/models.py:152-256): same constructor arguments, `common_step` /
`training_step` / `validation_step` / `test_step` / `configure_optimizers`, and the same logged names
(`{mode}_loss`, `{mode}_SISDR`, `{mode}_STFT`, `Input_SISDR`, `Input_STFT`).  What Lightning's Trainer does around
`training_step` under cfg/config.yaml:110-120 (automatic optimisation, `gradient_clip_val: 10.0`, fp32, scheduler
interval "step") is `fit_step`:

    zero_grad -> loss = training_step(batch) -> loss.backward() -> [all-reduce] clip-by-norm + AdamW -> scheduler.step()

Every piece of arithmetic is a kernel behind the C ABI: the network forward/backward (`TCNModel`: csrc/tcn.cu,
csrc/tcn_bwd.cu), the loss and its gradient (csrc/loss.cu), SI-SDR (csrc/loss.cu) and the flat-bucket optimiser
(csrc/optim.cu, whose `step()` also runs the one gradient all-reduce of a data-parallel step).  This file is host logic only.
Data-parallel use: one process per GPU, each rank feeds its own shard of the global batch; gradients are averaged in
`FusedAdamW.step()`, metrics logged with `sync_dist=True` in the reference are averaged over ranks here as well.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
from torch import Tensor, nn

from .losses import mrstft_loss, sisdr_loss
from .ops import causal_crop
from .optim import configure_optimizers as _configure_optimizers


class RemFX(nn.Module):
    """Drop-in for the Lightning module `remfx.models.RemFX` (cfg/model/*.yaml `_target_: remfx.models.RemFX`)."""

    def __init__(self, lr: float, lr_beta1: float, lr_beta2: float, lr_eps: float, lr_weight_decay: float, sample_rate: float,
                 network: nn.Module, max_steps: Optional[int] = None, gradient_clip_val: float = 10.0, process_group=None):
        super().__init__()
        self.lr = lr
        self.lr_beta1 = lr_beta1
        self.lr_beta2 = lr_beta2
        self.lr_eps = lr_eps
        self.lr_weight_decay = lr_weight_decay
        self.sample_rate = sample_rate
        self.model = network
        # the reference reads `self.trainer.max_steps` (cfg/config.yaml:113: 50000) and the Trainer's gradient_clip_val
        self.max_steps = max_steps
        self.gradient_clip_val = gradient_clip_val
        self.process_group = process_group
        self.log_train_audio = True
        self.output_str = "IN_SISDR,OUT_SISDR,IN_STFT,OUT_STFT\n"
        self.logged: Dict[str, Tensor] = {}  # last value of every `self.log(name, value)` call (0-d tensors; no host sync)
        self.compute_metrics = True          # the `no_grad` metric block of common_step (remfx/models.py:227-255)
        self._optim = None
        self._sched = None
        self.global_step = 0

    @property
    def device(self):
        return next(self.model.parameters()).device

    # ------------------------------------------------------------------ logging stand-in
    def log(self, name: str, value, sync_dist: bool = False, **_unused) -> None:
        v = value.detach() if isinstance(value, Tensor) else torch.as_tensor(float(value))
        if sync_dist:
            v = self._mean_over_ranks(v)
        self.logged[name] = v

    def _mean_over_ranks(self, v: Tensor) -> Tensor:
        import torch.distributed as dist

        if not dist.is_available() or not dist.is_initialized():
            return v
        world = dist.get_world_size(self.process_group)
        if world == 1:
            return v
        v = v.clone()
        dist.all_reduce(v, op=dist.ReduceOp.SUM, group=self.process_group)
        return v / world

    # ------------------------------------------------------------------ remfx/models.py:185-206
    def configure_optimizers(self):
        max_steps = self.max_steps
        if max_steps is None:
            trainer = getattr(self, "trainer", None)
            max_steps = getattr(trainer, "max_steps", None)
        if max_steps is None:
            raise ValueError("RemFX.configure_optimizers needs max_steps (the reference reads self.trainer.max_steps)")
        return _configure_optimizers(self.model, max_steps, lr=self.lr, lr_beta1=self.lr_beta1, lr_beta2=self.lr_beta2, lr_eps=self.lr_eps,
                                     lr_weight_decay=self.lr_weight_decay, gradient_clip_val=self.gradient_clip_val,
                                     process_group=self.process_group)

    # ------------------------------------------------------------------ remfx/models.py:208-256
    def training_step(self, batch, batch_idx: int = 0):
        return self.common_step(batch, batch_idx, mode="train")

    def validation_step(self, batch, batch_idx: int = 0):
        return self.common_step(batch, batch_idx, mode="valid")

    def test_step(self, batch, batch_idx: int = 0):
        return self.common_step(batch, batch_idx, mode="test")

    def common_step(self, batch, batch_idx: int = 0, mode: str = "train"):
        x, y, _, _ = batch
        if hasattr(self.model, "last_loss_terms"):
            self.model.last_loss_terms = None
        loss, output = self.model((x, y))
        terms = getattr(self.model, "last_loss_terms", None)   # the drop-in wrappers leave the loss kernel's 9 terms here
        target = y
        if output.shape[-1] < y.shape[-1]:
            target = causal_crop(y, output.shape[-1])
        self.log(f"{mode}_loss", loss)
        if self.compute_metrics:
            with torch.no_grad():
                out_d = output.detach()
                # SISDR is a loss (negative dB): logged negated, as the reference does
                self.log(f"{mode}_SISDR", -sisdr_loss(out_d, target), sync_dist=True)
                self.log("Input_SISDR", -sisdr_loss(x, y), sync_dist=True)
                # the MR-STFT value of (output, target) is a by-product of the loss kernels: the reference recomputes it here with six
                # more STFTs (remfx/models.py:236-245); a network that does not expose the terms gets the separate evaluation
                self.log(f"{mode}_STFT", terms[1] if terms is not None else mrstft_loss(out_d, target), sync_dist=True)
                self.log("Input_STFT", mrstft_loss(x, y), sync_dist=True)
        return loss

    # ------------------------------------------------------------------ what the Trainer does around training_step
    def fit_step(self, batch, batch_idx: int = 0, optimizer=None, scheduler=None) -> Tensor:
        """One optimisation step; returns the (detached, 0-d, device) training loss.  `optimizer` / `scheduler` default to the
        ones `configure_optimizers` builds (created on first use)."""
        if optimizer is None:
            if self._optim is None:
                cfg = self.configure_optimizers()
                self._optim, self._sched = cfg["optimizer"], cfg["lr_scheduler"]["scheduler"]
            optimizer, scheduler = self._optim, self._sched
        optimizer.zero_grad(set_to_none=True)  # torch's (and Lightning's) default; with FusedAdamW the next gradients land in one flat buffer
        loss = self.training_step(batch, batch_idx)
        loss.backward()
        optimizer.step()  # FusedAdamW: gradient all-reduce (if distributed) + clip-by-global-norm + AdamW
        if scheduler is not None:
            scheduler.step()
        self.global_step += 1
        return loss.detach()

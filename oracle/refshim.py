"""Import the UNCHANGED reference modules from /root/reference behind stubs.

TEST / MEASUREMENT INFRASTRUCTURE.  Reads ``/root/reference`` in the build container
and the unmodified, git-ignored copy ``oracle/_ref/`` (recipe: ``oracle/make_ref.py``)
where that does not exist (the GPU box).  Used by ``oracle/make_golden.py`` to generate
the committed fixtures under ``tests/golden/``, by the ``not gpu`` tests (skipped when
no reference tree is available) to pin the oracle restatements, and by ``bench.py``'s
reference arm / GPU-eager bar to time the unchanged reference modules.

The reference imports a number of third-party packages that are not installed
in this image.  None of them is on the numerical hot path
except auraloss, whose two loss classes are bound to the restatement in
``oracle/loss.py``.
"""
from __future__ import annotations

import os
import sys
import types

import torch
from torch import nn

_LOCAL_COPY = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")  # oracle/make_ref.py: unmodified copy that travels


def _pick_root() -> str:
    env = os.environ.get("REMFX_REFERENCE")
    if env:
        return env
    if os.path.isdir("/root/reference/remfx"):
        return "/root/reference"
    return _LOCAL_COPY


REF_ROOT = _pick_root()


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "remfx")) and os.path.isdir(
        os.path.join(REF_ROOT, "umx", "openunmix")
    )


def _mod(name: str, **attrs) -> types.ModuleType:
    m = sys.modules.get(name)
    if m is None:
        m = types.ModuleType(name)
        m.__path__ = []  # behave like a package so sub-imports work
        sys.modules[name] = m
        if "." in name:  # make `parent.child` attribute access work too
            parent, child = name.rsplit(".", 1)
            setattr(_mod(parent), child, m)
    for k, v in attrs.items():
        setattr(m, k, v)
    return m


class _LightningModule(nn.Module):
    """nn.Module with the handful of Lightning hooks the reference touches."""

    def log(self, *a, **k):
        return None

    def log_dict(self, *a, **k):
        return None

    @property
    def device(self):
        try:
            return next(self.parameters()).device
        except StopIteration:
            return torch.device("cpu")


class _Dummy:
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return None


_installed = False


def install() -> None:
    """Register the stub modules and put the reference on sys.path (idempotent)."""
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError(f"reference tree not found under {REF_ROOT}")
    from oracle import loss as _loss

    _mod(
        "pytorch_lightning",
        LightningModule=_LightningModule,
        LightningDataModule=object,
        Trainer=_Dummy,
        Callback=object,
        seed_everything=lambda s, **k: torch.manual_seed(s),
    )
    _mod("pytorch_lightning.utilities", rank_zero_only=lambda f: f)
    _mod("pytorch_lightning.utilities.rank_zero", rank_zero_only=lambda f: f)
    _mod("pytorch_lightning.callbacks", Callback=object)
    _mod("pytorch_lightning.loggers", CSVLogger=_Dummy, WandbLogger=_Dummy)
    _mod("pytorch_lightning.loggers.logger", Logger=object)
    _mod("omegaconf", DictConfig=dict, OmegaConf=_Dummy)
    _mod("torchmetrics", Accuracy=_Dummy)
    _mod("torchmetrics.classification", Accuracy=_Dummy, MultilabelF1Score=_Dummy)
    _mod("auraloss")
    _mod("auraloss.time", SISDRLoss=_loss.SISDRLoss)
    _mod("auraloss.freq", MultiResolutionSTFTLoss=_loss.MultiResolutionSTFTLoss)
    _mod("asteroid")
    _mod("asteroid.models", DCUNet=_Dummy)
    _mod("asteroid.models.dptnet", DPTNet=_Dummy)
    for name in ("hearbaseline", "hearbaseline.vggish", "hearbaseline.wav2vec2", "wav2clip_hear", "panns_hear"):
        _mod(name)
    _mod(
        "pedalboard",
        **{
            k: _Dummy
            for k in ("Pedalboard", "Chorus", "Reverb", "Compressor", "Phaser", "Delay", "Distortion", "Limiter")
        },
    )
    _mod("pyloudnorm", Meter=_Dummy)
    if "wandb" not in sys.modules:
        try:
            import wandb  # noqa: F401
        except Exception:
            _mod("wandb", Audio=_Dummy)
    for p in (REF_ROOT, os.path.join(REF_ROOT, "umx")):
        if p not in sys.path:
            sys.path.insert(0, p)
    _installed = True


def ref_modules():
    """Return a namespace with the reference classes on the hot path."""
    install()
    import remfx.models as rm  # type: ignore
    import remfx.tcn as rt  # type: ignore
    import remfx.classifier as rc  # type: ignore
    import remfx.utils as ru  # type: ignore
    from umx.openunmix import model as um  # type: ignore
    from umx.openunmix import transforms as ut  # type: ignore

    return types.SimpleNamespace(models=rm, tcn=rt, classifier=rc, utils=ru, umx_model=um, umx_transforms=ut)

"""Drop-in for `remfx.classifier.Cnn14` (remfx/classifier.py:134-284): same ctor kwargs, `forward(x) -> list of
(B, 1) probabilities`, same state_dict keys (92 tensors incl. `melspec.mel_scale.fb`, the unused `bn0.*`).
Sub-modules are parameter containers only; the math runs in libremfx_b200.so (csrc/cnn14.cu), eval mode
(`train=True` -- SpecAugment + dropout -- is not implemented on this inference path)."""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import torch
import torchaudio
from torch import Tensor, nn

from . import _lib


class _ConvBlockParams(nn.Module):
    """Parameter layout of remfx/classifier.py:236-267 (ConvBlock.__init__)."""

    def __init__(self, in_channels: int, out_channels: int):
        super().__init__()
        self.conv1 = nn.Conv2d(in_channels, out_channels, kernel_size=(3, 3), stride=(1, 1), padding=(1, 1), bias=False)
        self.conv2 = nn.Conv2d(out_channels, out_channels, kernel_size=(3, 3), stride=(1, 1), padding=(1, 1), bias=False)
        self.bn1 = nn.BatchNorm2d(out_channels)
        self.bn2 = nn.BatchNorm2d(out_channels)
        for conv in (self.conv1, self.conv2):
            nn.init.xavier_uniform_(conv.weight)


class Cnn14(nn.Module):
    def __init__(self, num_classes: int, sample_rate: float, model_sample_rate: float, n_fft: int = 1024, hop_length: int = 256,
                 n_mels: int = 128, specaugment: bool = False):
        super().__init__()
        if sample_rate != model_sample_rate:
            raise ValueError("remfx_b200.Cnn14 needs sample_rate == model_sample_rate (no resampler on this path; RemFx uses 48 kHz for both)")
        self.num_classes = num_classes
        self.n_fft = n_fft
        self.hop_length = hop_length
        self.n_mels = n_mels
        self.sample_rate = sample_rate
        self.model_sample_rate = model_sample_rate
        self.specaugment = specaugment
        self.register_buffer("window", torch.hann_window(n_fft))
        self.melspec = torchaudio.transforms.MelSpectrogram(model_sample_rate, n_fft, hop_length=hop_length, n_mels=n_mels)
        self.bn0 = nn.BatchNorm2d(n_mels)  # allocated but unused by the reference as well
        chans = [1, 64, 128, 256, 512, 1024, 2048]
        for i in range(6):
            setattr(self, f"conv_block{i + 1}", _ConvBlockParams(chans[i], chans[i + 1]))
        self.fc1 = nn.Linear(2048, 2048, bias=True)
        nn.init.xavier_uniform_(self.fc1.weight)
        self.fc1.bias.data.fill_(0.0)
        self.heads = nn.ModuleList([nn.Linear(2048, 1, bias=True) for _ in range(num_classes)])
        self._handle: Optional[C.c_void_p] = None
        self._stamp = None
        self._ws: Optional[Tensor] = None

    def _sync(self, device) -> C.c_void_p:
        tensors = {k: v for k, v in self.state_dict(keep_vars=True).items() if v.dtype == torch.float32 and not k.startswith("bn0.") and k != "window"}
        stamp = (str(device),) + tuple((k, t.data_ptr(), t._version) for k, t in tensors.items())
        L = _lib.lib()
        if self._handle is not None and stamp == self._stamp:
            return self._handle
        if self._handle is None:
            cfg = _lib.Cnn14Config(self.num_classes, self.n_fft, self.hop_length, self.n_mels)
            h = C.c_void_p()
            _lib.check(L.rfx_cnn14_create(C.byref(cfg), C.byref(h)), "rfx_cnn14_create")
            self._handle = h
        stream = _lib.cur_stream()
        for k, t in tensors.items():
            if t.device != device:
                raise _lib.RfxError(f"parameter {k} is on {t.device}, input on {device}: call .to(device) first")
            tc = t.detach().contiguous()
            _lib.check(L.rfx_cnn14_load_param(self._handle, k.encode(), tc.data_ptr(), tc.numel(), stream), f"load {k}")
        _lib.check(L.rfx_cnn14_finalize(self._handle, stream), "rfx_cnn14_finalize")
        self._stamp = stamp
        return self._handle

    def __del__(self):
        h = self.__dict__.get("_handle")
        if h is not None:
            self.__dict__["_handle"] = None
            try:
                _lib.lib().rfx_cnn14_destroy(h)
            except Exception:
                pass

    def probs_and_logits(self, x: Tensor):
        """(B, 1, T) or (B, T) -> ((B, K) probabilities, (B, K) logits)."""
        if x.dim() == 3 and x.shape[1] == 1:
            x = x[:, 0]
        if x.dim() != 2:
            raise ValueError(f"expected (batch, 1, time) or (batch, time), got {tuple(x.shape)}")
        _lib.require_device(x)
        if x.dtype != torch.float32:
            raise ValueError("expected float32 audio")
        x = x.contiguous()
        B, T = x.shape
        L = _lib.lib()
        with torch.cuda.device(x.device):
            h = self._sync(x.device)
            need = L.rfx_cnn14_workspace_bytes(h, B, T)
            if self._ws is None or self._ws.numel() < need or self._ws.device != x.device:
                self._ws = torch.empty(need, dtype=torch.uint8, device=x.device)
            probs = torch.empty(B, self.num_classes, dtype=torch.float32, device=x.device)
            logits = torch.empty(B, self.num_classes, dtype=torch.float32, device=x.device)
            rc = L.rfx_cnn14_forward(h, x.data_ptr(), B, T, probs.data_ptr(), logits.data_ptr(), self._ws.data_ptr(), self._ws.numel(),
                                     _lib.cur_stream())
            _lib.check(rc, "rfx_cnn14_forward")
        return probs, logits

    def forward(self, x: Tensor, train: bool = False) -> List[Tensor]:
        """List of `num_classes` (B, 1) sigmoid outputs, as the reference returns (classifier.py:229-233)."""
        if train:
            raise NotImplementedError("remfx_b200.Cnn14 implements the inference path only (train=False)")
        probs, _ = self.probs_and_logits(x)
        return [probs[:, k : k + 1] for k in range(self.num_classes)]

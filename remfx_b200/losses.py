"""RemFx loss on the GPU: MRSTFT(out, target) + 100 * L1(out, target) (remfx/models.py:299,320,385).

The fused kernels never materialise a spectrogram in HBM.  `remfx_loss` is differentiable with respect to `out`
(`rfx_remfx_loss_backward`: the first link of the training step; the TCN's and Hybrid Demucs' backward kernels continue
from its gradient -- csrc/tcn_bwd.cu, csrc/hdemucs_bwd.cu).
See csrc/loss.cu and oracle/loss.py (auraloss restatement, parity unpinned).
"""
from __future__ import annotations

import torch
from torch import Tensor

from . import _lib
from .ops import padded_window

_WIN_CACHE: dict = {}
_RES = ((1024, 600), (2048, 1200), (512, 240))


def _windows(device):
    key = str(device)
    if key not in _WIN_CACHE:
        _WIN_CACHE[key] = [padded_window(torch.hann_window(w, device=device), n) for n, w in _RES]
    return _WIN_CACHE[key]


def remfx_loss_terms(out: Tensor, target: Tensor, l1_weight: float = 100.0) -> Tensor:
    """out, target: (B, C, T) or (B, T) fp32 CUDA (last-dim contiguous views are fine) -> 9-float device tensor
    [loss, mrstft, mean|d|, sc_1024, lm_1024, sc_2048, lm_2048, sc_512, lm_512]."""
    _lib.require_device(out)
    _lib.require_device(target)
    if out.shape != target.shape:
        raise ValueError(f"shape mismatch {tuple(out.shape)} vs {tuple(target.shape)}")
    T = out.shape[-1]
    o2 = out.reshape(-1, T)
    t2 = target.reshape(-1, T)
    if o2.stride(-1) != 1:
        o2 = o2.contiguous()
    if t2.stride(-1) != 1:
        t2 = t2.contiguous()
    if o2.dtype != torch.float32 or t2.dtype != torch.float32:
        raise ValueError("expected float32")
    B = o2.shape[0]
    L = _lib.lib()
    with torch.cuda.device(out.device):
        ws = torch.empty(L.rfx_loss_workspace_bytes(B, T), dtype=torch.uint8, device=out.device)
        res = torch.empty(9, dtype=torch.float32, device=out.device)
        w = _windows(out.device)
        rc = L.rfx_remfx_loss(o2.data_ptr(), o2.stride(0) if B > 1 else T, t2.data_ptr(), t2.stride(0) if B > 1 else T, B, T,
                              w[0].data_ptr(), w[1].data_ptr(), w[2].data_ptr(), float(l1_weight), res.data_ptr(),
                              ws.data_ptr(), ws.numel(), _lib.cur_stream())
        _lib.check(rc, "rfx_remfx_loss")
    return res


def _rows2(t: Tensor):
    T = t.shape[-1]
    t2 = t.reshape(-1, T)
    if t2.stride(-1) != 1:
        t2 = t2.contiguous()
    return t2, (t2.stride(0) if t2.shape[0] > 1 else T)


class _RemfxLossFn(torch.autograd.Function):
    """loss = MRSTFT(out, target) + l1_weight * L1(out, target) with d loss / d out from rfx_remfx_loss_backward."""

    @staticmethod
    def forward(ctx, out: Tensor, target: Tensor, l1_weight: float):
        o2, obs = _rows2(out.detach())
        t2, tbs = _rows2(target.detach())
        B, T = o2.shape
        L = _lib.lib()
        with torch.cuda.device(out.device):
            ws = torch.empty(L.rfx_loss_workspace_bytes(B, T), dtype=torch.uint8, device=out.device)
            res = torch.empty(9, dtype=torch.float32, device=out.device)
            w = _windows(out.device)
            rc = L.rfx_remfx_loss(o2.data_ptr(), obs, t2.data_ptr(), tbs, B, T, w[0].data_ptr(), w[1].data_ptr(), w[2].data_ptr(),
                                  float(l1_weight), res.data_ptr(), ws.data_ptr(), ws.numel(), _lib.cur_stream())
            _lib.check(rc, "rfx_remfx_loss")
        ctx.save_for_backward(o2, t2, ws)
        ctx.meta = (obs, tbs, B, T, float(l1_weight), out.shape)
        ctx.mark_non_differentiable(res)
        return res[0].clone(), res

    @staticmethod
    def backward(ctx, grad_loss: Tensor, _grad_terms=None):
        o2, t2, ws = ctx.saved_tensors
        obs, tbs, B, T, l1w, shape = ctx.meta
        L = _lib.lib()
        with torch.cuda.device(o2.device):
            g = torch.empty(B, T, dtype=torch.float32, device=o2.device)
            gl = grad_loss.detach().to(torch.float32).reshape(1).contiguous()
            w = _windows(o2.device)
            rc = L.rfx_remfx_loss_backward(o2.data_ptr(), obs, t2.data_ptr(), tbs, B, T, w[0].data_ptr(), w[1].data_ptr(), w[2].data_ptr(),
                                           l1w, gl.data_ptr(), g.data_ptr(), T, ws.data_ptr(), ws.numel(), _lib.cur_stream())
            _lib.check(rc, "rfx_remfx_loss_backward")
        return g.reshape(shape), None, None


def remfx_loss_with_terms(out: Tensor, target: Tensor):
    """(loss, terms): the 0-d loss the reference wrappers return (differentiable with respect to `out`) and the detached 9-vector
    of `remfx_loss_terms`.  terms[1] is the MR-STFT value of (out, target) -- exactly what the reference's `no_grad` metric block
    recomputes with six more STFTs (remfx/models.py:236-245); the training harness reuses it instead (row N3)."""
    if out.requires_grad and torch.is_grad_enabled():
        _lib.require_device(out)
        _lib.require_device(target)
        if out.shape != target.shape:
            raise ValueError(f"shape mismatch {tuple(out.shape)} vs {tuple(target.shape)}")
        if out.dtype != torch.float32 or target.dtype != torch.float32:
            raise ValueError("expected float32")
        loss, terms = _RemfxLossFn.apply(out, target, 100.0)
        return loss, terms.detach()
    terms = remfx_loss_terms(out, target)
    return terms[0], terms


def remfx_loss(out: Tensor, target: Tensor) -> Tensor:
    """0-d loss tensor, as the reference wrappers return; differentiable with respect to `out`."""
    return remfx_loss_with_terms(out, target)[0]


def mrstft_loss(out: Tensor, target: Tensor) -> Tensor:
    return remfx_loss_terms(out, target)[1]


def _rows(t: Tensor):
    T = t.shape[-1]
    t2 = t.reshape(-1, T)
    if t2.stride(-1) != 1:
        t2 = t2.contiguous()
    return t2, (t2.stride(0) if t2.shape[0] > 1 else T)


def sisdr_loss(inp: Tensor, target: Tensor) -> Tensor:
    """auraloss SISDRLoss() value (negative SI-SDR in dB, batch mean) as a 0-d device tensor (remfx/models.py:41,173)."""
    _lib.require_device(inp)
    _lib.require_device(target)
    if inp.shape != target.shape:
        raise ValueError(f"shape mismatch {tuple(inp.shape)} vs {tuple(target.shape)}")
    if inp.dtype != torch.float32 or target.dtype != torch.float32:
        raise ValueError("expected float32")
    a, abs_ = _rows(inp)
    b, bbs = _rows(target)
    B, T = a.shape
    L = _lib.lib()
    with torch.cuda.device(inp.device):
        ws = torch.empty(L.rfx_sisdr_workspace_bytes(B), dtype=torch.uint8, device=inp.device)
        res = torch.empty(1, dtype=torch.float32, device=inp.device)
        rc = L.rfx_sisdr_loss(a.data_ptr(), abs_, b.data_ptr(), bbs, B, T, res.data_ptr(), ws.data_ptr(), ws.numel(), _lib.cur_stream())
        _lib.check(rc, "rfx_sisdr_loss")
    return res[0]

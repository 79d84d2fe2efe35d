"""Stand-alone operators of the hot path, as thin wrappers over the C ABI.

Mirrors of the reference's small building blocks:
  stft / istft        -- umx/openunmix/transforms.py:89-120, 164-181 (TorchSTFT / TorchISTFT)
  complex_norm        -- umx/openunmix/transforms.py:198-216
  spectrogram         -- remfx/utils.py:138-159
  center_crop / causal_crop -- remfx/utils.py:202-211 (pure views, no kernel)
Outputs follow the reference's shapes (…, bins, frames[, 2]); internally the kernels produce
frame-major data, so these wrappers return permuted *views* (no copy).
"""
from __future__ import annotations

import torch

from . import _lib


def _prep(x: torch.Tensor) -> torch.Tensor:
    _lib.require_device(x)
    if x.dtype != torch.float32:
        raise ValueError(f"expected float32, got {x.dtype}")
    return x.contiguous()


def padded_window(window: torch.Tensor, n_fft: int) -> torch.Tensor:
    """Zero-pad (centred) a short window to n_fft taps, as torch.stft does for win_length < n_fft."""
    w = window.to(torch.float32)
    if w.numel() == n_fft:
        return w.contiguous()
    left = (n_fft - w.numel()) // 2
    out = torch.zeros(n_fft, dtype=torch.float32, device=w.device)
    out[left : left + w.numel()] = w
    return out


_MODES = {"complex": 0, "mag": 2, "power": 3, "mag_clamp": 4, "mag_pow": 5}


def stft_raw(x: torch.Tensor, n_fft: int, hop: int, window: torch.Tensor, normalized: bool = False, mode: str = "complex",
             alpha: float = 1.0, want_complex: bool = True):
    """x: (N, T) -> (Z, A): Z (N, F, bins, 2) frame-major complex or None, A (N, F, bins) or None."""
    x = _prep(x)
    if x.dim() != 2:
        raise ValueError("stft_raw expects (N, T)")
    N, T = x.shape
    F, bins = T // hop + 1, n_fft // 2 + 1
    w = padded_window(_prep(window), n_fft)
    Z = torch.empty(N, F, bins, 2, dtype=torch.float32, device=x.device) if (want_complex or mode == "complex") else None
    A = torch.empty(N, F, bins, dtype=torch.float32, device=x.device) if mode != "complex" else None
    rc = _lib.lib().rfx_stft(_lib.ptr(x), N, T, n_fft, hop, _lib.ptr(w), int(normalized), _MODES[mode], float(alpha),
                             _lib.ptr(Z), _lib.ptr(A), _lib.cur_stream())
    _lib.check(rc, "rfx_stft")
    return Z, A


def stft(x: torch.Tensor, n_fft: int = 4096, n_hop: int = 1024, window: torch.Tensor | None = None) -> torch.Tensor:
    """TorchSTFT.forward: (B, C, T) -> (B, C, bins, frames, 2); center=True, reflect padding."""
    if x.dim() != 3:
        raise ValueError("stft expects (nb_samples, nb_channels, nb_timesteps)")
    B, Cn, T = x.shape
    if window is None:
        window = torch.hann_window(n_fft, device=x.device)
    Z, _ = stft_raw(x.reshape(B * Cn, T), n_fft, n_hop, window)
    return Z.view(B, Cn, Z.shape[1], Z.shape[2], 2).permute(0, 1, 3, 2, 4)


def istft(X: torch.Tensor, n_fft: int = 4096, n_hop: int = 1024, window: torch.Tensor | None = None, length: int | None = None,
          normalized: bool = False) -> torch.Tensor:
    """TorchISTFT.forward: (..., bins, frames, 2) -> (..., length); center=True."""
    _lib.require_device(X)
    shape = X.shape
    bins, F = shape[-3], shape[-2]
    Zf = X.reshape(-1, bins, F, 2).permute(0, 2, 1, 3).contiguous()  # frame-major
    N = Zf.shape[0]
    if window is None:
        window = torch.hann_window(n_fft, device=X.device)
    if length is None:
        length = n_hop * (F - 1)
    w = padded_window(_prep(window), n_fft)
    out = torch.empty(N, length, dtype=torch.float32, device=X.device)
    rc = _lib.lib().rfx_istft(_lib.ptr(Zf), 0, N, F, n_fft, n_hop, _lib.ptr(w), int(normalized), int(length), _lib.ptr(out),
                              _lib.cur_stream())
    _lib.check(rc, "rfx_istft")
    return out.reshape(shape[:-3] + (length,))


def complex_norm(spec: torch.Tensor, mono: bool = False) -> torch.Tensor:
    """ComplexNorm.forward on a (..., 2) tensor (cheap elementwise; the fused path never materialises it)."""
    mag = torch.linalg.vector_norm(spec, dim=-1)
    return mag.mean(1, keepdim=True) if mono else mag


def spectrogram(x: torch.Tensor, window: torch.Tensor, n_fft: int, hop_length: int, alpha: float) -> torch.Tensor:
    """remfx.utils.spectrogram: (B, C, T) -> (B, C, bins, frames) = (|STFT| + 1e-8) ** alpha."""
    if x.dim() != 3:
        raise ValueError("spectrogram expects (bs, chs, samp)")
    B, Cn, T = x.shape
    _, A = stft_raw(x.reshape(B * Cn, T), n_fft, hop_length, window, mode="mag_pow", alpha=alpha, want_complex=False)
    return A.view(B, Cn, A.shape[1], A.shape[2]).permute(0, 1, 3, 2)


def center_crop(x: torch.Tensor, length: int) -> torch.Tensor:
    start = (x.shape[-1] - length) // 2
    return x[..., start : start + length]


def causal_crop(x: torch.Tensor, length: int) -> torch.Tensor:
    stop = x.shape[-1] - 1
    return x[..., stop - length : stop]


_ACTS = {None: 0, "none": 0, "tanh": 1, "relu": 2, "sigmoid": 3}


def linear(A: torch.Tensor, W: torch.Tensor, s1=None, t1=None, s2=None, t2=None, act=None, impl: str = "tc") -> torch.Tensor:
    """act(((A @ W.T) * s1 + t1) * s2 + t2) with A (M, K) [row stride multiple of 4], W (N, K)."""
    A = _prep(A)
    W = _prep(W)
    M, K = A.shape
    N = W.shape[0]
    if W.shape[1] != K:
        raise ValueError("linear: K mismatch")
    L = _lib.lib()
    Ain, lda = A, K
    Cout = torch.empty(M, N, dtype=torch.float32, device=A.device)
    scratch = None
    if impl == "tc":
        scratch = torch.empty(L.rfx_gemm_scratch_bytes(M, N, K), dtype=torch.uint8, device=A.device)
    vecs = [None if v is None else _prep(v) for v in (s1, t1, s2, t2)]
    rc = L.rfx_gemm(0 if impl == "tc" else 1, _lib.ptr(Ain), lda, M, _lib.ptr(W), N, K, _lib.ptr(Cout), N, *[_lib.ptr(v) for v in vecs],
                    _ACTS[act], _lib.ptr(scratch), _lib.cur_stream())
    _lib.check(rc, "rfx_gemm")
    return Cout


def lstm_layer(G: torch.Tensor, Whh: torch.Tensor, B: int, F: int, impl: str = "mma", slots: int = 0) -> torch.Tensor:
    """Recurrent half of one bidirectional LSTM layer.  G: (B*F, 8H) input projections (+biases),
    Whh: (2, 4H, H) -> (B*F, 2H).  impl: "mma" (mma.sync bf16x3, default), "ffma" (fp32) or "tc" (tcgen05, H = 256).
    slots: batch slots per cluster (0 = automatic; at most 8, or 16 for "tc")."""
    _lib.check(_lib.lib().rfx_lstm_set_impl({"mma": 0, "ffma": 1, "tc": 2}[impl]), "rfx_lstm_set_impl")
    try:
        return _lstm_layer(G, Whh, B, F, slots)
    finally:
        _lib.lib().rfx_lstm_set_impl(0)


def _lstm_layer(G: torch.Tensor, Whh: torch.Tensor, B: int, F: int, slots: int = 0) -> torch.Tensor:
    G = _prep(G)
    Whh = _prep(Whh)
    H = Whh.shape[-1]
    out = torch.empty(B * F, 2 * H, dtype=torch.float32, device=G.device)
    rc = _lib.lib().rfx_lstm_layer_slots(_lib.ptr(G), _lib.ptr(Whh), _lib.ptr(out), 2 * H, B, F, H, int(slots), _lib.cur_stream())
    _lib.check(rc, "rfx_lstm_layer_slots")
    return out

"""ctypes binding of libremfx_b200.so (the C ABI declared in include/remfx_b200.h).

The product path has NO fallback: if the CUDA library is missing or the device is not a
Blackwell (sm_100) part, the ops raise instead of silently computing elsewhere.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libremfx_b200.so")

_lock = threading.Lock()
_lib = None


class RfxError(RuntimeError):
    pass


class TcnConfig(C.Structure):
    _fields_ = [("ninputs", C.c_int), ("noutputs", C.c_int), ("nblocks", C.c_int), ("channel_width", C.c_int), ("kernel_size", C.c_int),
                ("stack_size", C.c_int), ("dilation_growth", C.c_int), ("causal", C.c_int)]


class Cnn14Config(C.Structure):
    _fields_ = [("num_classes", C.c_int), ("n_fft", C.c_int), ("hop", C.c_int), ("n_mels", C.c_int)]


class HDemucsConfig(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("audio_channels", "n_sources", "channels", "growth", "nfft", "depth", "kernel_size", "stride",
                                        "time_stride", "context", "context_enc", "norm_starts", "norm_groups", "dconv_depth", "dconv_comp",
                                        "dconv_attn", "dconv_lstm")] + [("freq_emb_weight", C.c_float), ("freq_emb_scale", C.c_float)]


class UmxConfig(C.Structure):
    _fields_ = [("n_fft", C.c_int), ("hop", C.c_int), ("hidden", C.c_int), ("nb_layers", C.c_int), ("gemm_impl", C.c_int)]


_f32p = C.c_void_p  # device pointers travel as integers
_SIGNATURES = {
    "rfx_abi_version": (C.c_int, []),
    "rfx_last_error": (C.c_char_p, []),
    "rfx_device_supported": (C.c_int, []),
    "rfx_stft": (C.c_int, [_f32p, C.c_int, C.c_int, C.c_int, C.c_int, _f32p, C.c_int, C.c_int, C.c_float, _f32p, _f32p, C.c_void_p]),
    "rfx_istft": (C.c_int, [_f32p, _f32p, C.c_int, C.c_int, C.c_int, C.c_int, _f32p, C.c_int, C.c_int, _f32p, C.c_void_p]),
    "rfx_gemm_scratch_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "rfx_gemm": (C.c_int, [C.c_int, _f32p, C.c_int, C.c_int, _f32p, C.c_int, C.c_int, _f32p, C.c_int, _f32p, _f32p, _f32p, _f32p,
                           C.c_int, C.c_void_p, C.c_void_p]),
    "rfx_set_matmul_precision": (C.c_int, [C.c_int]),
    "rfx_get_matmul_precision": (C.c_int, []),
    "rfx_lstm_info": (C.c_int, [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "rfx_lstm_set_impl": (C.c_int, [C.c_int]),
    "rfx_lstm_layer": (C.c_int, [_f32p, _f32p, _f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "rfx_lstm_layer_slots": (C.c_int, [_f32p, _f32p, _f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "rfx_umx_create": (C.c_int, [C.POINTER(UmxConfig), C.POINTER(C.c_void_p)]),
    "rfx_umx_destroy": (None, [C.c_void_p]),
    "rfx_umx_load_param": (C.c_int, [C.c_void_p, C.c_char_p, _f32p, C.c_int64, C.c_void_p]),
    "rfx_umx_finalize": (C.c_int, [C.c_void_p, C.c_void_p]),
    "rfx_umx_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int, C.c_int]),
    "rfx_umx_sample": (C.c_int, [C.c_void_p, _f32p, C.c_int, C.c_int, _f32p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "rfx_umx_sample_host": (C.c_int, [C.c_void_p, _f32p, C.c_int, C.c_int, _f32p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "rfx_umx_submit_host": (C.c_int, [C.c_void_p, C.c_int, _f32p, C.c_int, C.c_int, _f32p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "rfx_umx_wait_host": (C.c_int, [C.c_void_p, C.c_int]),
    "rfx_umx_pipe_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int, C.c_int]),
    "rfx_umx_pipe_depth": (C.c_int, [C.c_void_p]),
    "rfx_umx_pipe_info": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "rfx_umx_pipe_push": (C.c_int, [C.c_void_p, _f32p, C.c_int, C.c_int, C.c_int, _f32p, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p,
                                    C.POINTER(C.c_longlong)]),
    "rfx_umx_pipe_flush": (C.c_int, [C.c_void_p, C.c_void_p]),
    "rfx_umx_pipe_wait": (C.c_int, [C.c_void_p, C.c_longlong]),
    "rfx_umx_pipe_query": (C.c_int, [C.c_void_p, C.c_longlong, C.POINTER(C.c_int)]),
    "rfx_umx_pipe_stream_wait": (C.c_int, [C.c_void_p, C.c_longlong, C.c_void_p]),
    "rfx_umx_pipe_set_profiling": (C.c_int, [C.c_void_p, C.c_int]),
    "rfx_umx_pipe_rec_times": (C.c_int, [C.c_void_p, C.POINTER(C.c_float), C.c_int, C.POINTER(C.c_int)]),
    "rfx_umx_launches_per_call": (C.c_int, [C.c_void_p]),
    "rfx_umx_set_profiling": (C.c_int, [C.c_void_p, C.c_int]),
    "rfx_umx_stage_times": (C.c_int, [C.c_void_p, C.POINTER(C.c_float), C.c_int, C.POINTER(C.c_int)]),
    "rfx_umx_debug_tap": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, _f32p, C.POINTER(C.c_int), C.c_void_p]),
    "rfx_tcn_create": (C.c_int, [C.POINTER(TcnConfig), C.POINTER(C.c_void_p)]),
    "rfx_tcn_destroy": (None, [C.c_void_p]),
    "rfx_tcn_load_param": (C.c_int, [C.c_void_p, C.c_char_p, _f32p, C.c_int64, C.c_void_p]),
    "rfx_tcn_finalize": (C.c_int, [C.c_void_p, C.c_void_p]),
    "rfx_tcn_out_length": (C.c_longlong, [C.c_void_p, C.c_longlong]),
    "rfx_tcn_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int, C.c_longlong]),
    "rfx_tcn_forward": (C.c_int, [C.c_void_p, _f32p, C.c_int, C.c_longlong, _f32p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "rfx_tcn_launches_per_call": (C.c_int, [C.c_void_p]),
    "rfx_tcn_train_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int, C.c_longlong]),
    "rfx_tcn_forward_train": (C.c_int, [C.c_void_p, _f32p, C.c_int, C.c_longlong, _f32p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "rfx_tcn_backward": (C.c_int, [C.c_void_p, _f32p, _f32p, _f32p, C.c_int, C.c_longlong, C.POINTER(C.c_char_p), C.POINTER(C.c_void_p),
                                   C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]),
    "rfx_tcn_backward_launches_per_call": (C.c_int, [C.c_void_p]),
    "rfx_tcn_set_wgrad_impl": (C.c_int, [C.c_int]),
    "rfx_cnn14_create": (C.c_int, [C.POINTER(Cnn14Config), C.POINTER(C.c_void_p)]),
    "rfx_cnn14_destroy": (None, [C.c_void_p]),
    "rfx_cnn14_load_param": (C.c_int, [C.c_void_p, C.c_char_p, _f32p, C.c_int64, C.c_void_p]),
    "rfx_cnn14_finalize": (C.c_int, [C.c_void_p, C.c_void_p]),
    "rfx_cnn14_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int, C.c_int]),
    "rfx_cnn14_forward": (C.c_int, [C.c_void_p, _f32p, C.c_int, C.c_int, _f32p, _f32p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "rfx_hdemucs_create": (C.c_int, [C.POINTER(HDemucsConfig), C.POINTER(C.c_void_p)]),
    "rfx_hdemucs_destroy": (None, [C.c_void_p]),
    "rfx_hdemucs_load_param": (C.c_int, [C.c_void_p, C.c_char_p, _f32p, C.c_int64, C.c_void_p]),
    "rfx_hdemucs_finalize": (C.c_int, [C.c_void_p, C.c_void_p]),
    "rfx_hdemucs_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int, C.c_int]),
    "rfx_hdemucs_forward": (C.c_int, [C.c_void_p, _f32p, C.c_int, C.c_int, _f32p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "rfx_hdemucs_launches_per_call": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "rfx_hdemucs_set_taps": (C.c_int, [C.c_void_p, C.c_int]),
    "rfx_hdemucs_tap": (C.c_int, [C.c_void_p, C.c_char_p, _f32p, C.c_int64, C.POINTER(C.c_int), C.c_void_p]),
    "rfx_umx_train_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int, C.c_int]),
    "rfx_umx_train_prepare": (C.c_int, [C.c_void_p, C.c_void_p]),
    "rfx_umx_forward_train": (C.c_int, [C.c_void_p, _f32p, C.c_int, C.c_int, _f32p, C.c_void_p, C.c_size_t, _f32p, C.c_int, C.c_float, _f32p,
                                        C.c_void_p]),
    "rfx_umx_backward": (C.c_int, [C.c_void_p, _f32p, _f32p, C.c_int, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_void_p), C.c_int,
                                   C.c_void_p, C.c_size_t, C.c_void_p]),
    "rfx_hdemucs_train_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int, C.c_int]),
    "rfx_hdemucs_forward_train": (C.c_int, [C.c_void_p, _f32p, C.c_int, C.c_int, _f32p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "rfx_hdemucs_backward": (C.c_int, [C.c_void_p, _f32p, _f32p, C.c_int, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_void_p), C.c_int,
                                       C.c_void_p, C.c_size_t, C.c_void_p]),
    "rfx_hdemucs_set_wgrad_impl": (C.c_int, [C.c_int]),
    "rfx_hdemucs_set_grads_prezeroed": (C.c_int, [C.c_void_p, C.c_int]),
    "rfx_hdemucs_grad_tap": (C.c_int, [C.c_void_p, C.c_char_p, _f32p, C.c_int64, C.POINTER(C.c_int), C.c_void_p]),
    "rfx_hdemucs_inject_grad": (C.c_int, [C.c_void_p, C.c_char_p, _f32p]),
    "rfx_loss_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "rfx_remfx_loss": (C.c_int, [_f32p, C.c_longlong, _f32p, C.c_longlong, C.c_int, C.c_int, _f32p, _f32p, _f32p, C.c_float, _f32p,
                                 C.c_void_p, C.c_size_t, C.c_void_p]),
    "rfx_remfx_loss_backward": (C.c_int, [_f32p, C.c_longlong, _f32p, C.c_longlong, C.c_int, C.c_int, _f32p, _f32p, _f32p, C.c_float, _f32p,
                                          _f32p, C.c_longlong, C.c_void_p, C.c_size_t, C.c_void_p]),
    "rfx_sisdr_workspace_bytes": (C.c_size_t, [C.c_int]),
    "rfx_sisdr_loss": (C.c_int, [_f32p, C.c_longlong, _f32p, C.c_longlong, C.c_int, C.c_int, _f32p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "rfx_wav_info": (C.c_int, [C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_longlong), C.POINTER(C.c_int),
                               C.POINTER(C.c_int)]),
    "rfx_ingest_wav_batch": (C.c_int, [C.POINTER(C.c_char_p), C.c_int, C.c_void_p, C.c_longlong, C.c_int, C.POINTER(C.c_longlong),
                                       C.POINTER(C.c_int)]),
    "rfx_optim_workspace_bytes": (C.c_size_t, []),
    "rfx_grad_sumsq": (C.c_int, [_f32p, C.c_longlong, C.c_void_p, C.c_int, C.c_void_p]),
    "rfx_adamw_step": (C.c_int, [_f32p, _f32p, _f32p, _f32p, C.c_longlong, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float,
                                 C.c_int, C.c_float, C.c_float, C.c_void_p, _f32p, C.c_void_p]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)


def lib() -> C.CDLL:
    """Load (once) and return the shared library; raises RfxError when it has not been built."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise RfxError(
                    f"{LIB_PATH} not found: build it with `python -m remfx_b200.build` "
                    "(remfx_b200 has no CPU or eager-PyTorch fallback)"
                )
            handle = C.CDLL(LIB_PATH)
            for name, (res, args) in _SIGNATURES.items():
                fn = getattr(handle, name)
                fn.restype = res
                fn.argtypes = args
            if handle.rfx_abi_version() != 1:
                raise RfxError("libremfx_b200.so ABI version mismatch; rebuild")
            _lib = handle
    return _lib


def set_precision(mode: str) -> None:
    """Process-wide matmul precision of the tensor-core kernels: "fp32" (bf16x3, parity mode, default) or "bf16" (single pass)."""
    if mode not in ("fp32", "bf16"):
        raise ValueError("precision must be 'fp32' (parity, bf16x3) or 'bf16' (fast, single pass)")
    check(lib().rfx_set_matmul_precision(1 if mode == "bf16" else 0), "rfx_set_matmul_precision")


def get_precision() -> str:
    return "bf16" if lib().rfx_get_matmul_precision() == 1 else "fp32"


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().rfx_last_error().decode("utf-8", "replace")
        exc = ValueError if rc == 2 else RfxError
        raise exc(f"{what}: {msg}" if what else msg)


def require_device(t) -> None:
    """Raise unless tensor `t` lives on a supported CUDA device."""
    if not t.is_cuda:
        raise RfxError("remfx_b200 kernels need CUDA tensors on a B200 (sm_100a); got a CPU tensor and there is no CPU fallback")


def ptr(t) -> int:
    return 0 if t is None else t.data_ptr()


def cur_stream() -> int:
    import torch

    return torch.cuda.current_stream().cuda_stream

"""remfx_b200 -- Blackwell-native (sm_100a) compute path for RemFx audio-effect removal.

Public surface mirrors the reference's plug-in boundary (remfx/models.py:259-390):
`remfx_b200.models.{OpenUnmixModel, ...}` with `forward((x, target)) -> (loss, out)` and
`sample(x) -> out`; `remfx_b200.ops` for the stand-alone STFT/iSTFT/crop helpers.  All math runs
in hand-written CUDA kernels behind the C ABI in include/remfx_b200.h (libremfx_b200.so).
"""
__version__ = "0.1.0"


def set_precision(mode: str) -> None:
    """"fp32" (default: bf16x3 products, fp32-grade parity) or "bf16" (single-pass tensor-core products, the fast mode)."""
    from . import _lib

    _lib.set_precision(mode)


def get_precision() -> str:
    from . import _lib

    return _lib.get_precision()

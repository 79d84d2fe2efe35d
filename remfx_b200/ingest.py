"""Batch ingest: pre-rendered chunk directories -> pinned (B, 1, T) batches for the device pipeline.

Mirrors the item contract of `remfx.datasets.EffectDataset.__getitem__` (remfx/datasets.py:461-468) on an already rendered
`proc_root` (`<root>/<idx>/{input.wav,target.wav,dry_effects.pt,wet_effects.pt}`, written at remfx/datasets.py:197-200,447-450):
an item is `(input (1, T), target (1, T), dry_effects (5,), wet_effects (5,))`.  The reference decodes item by item in DataLoader
worker processes and collates; here a whole batch is decoded by the native thread pool (`rfx_ingest_wav_batch`) straight into
one pinned (B, 1, T) buffer per signal, which `UmxPipeline.push` / `sample_host` copy to the device themselves.
Rendering the chunks (effects, loudness normalisation) stays out of scope.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Iterator, List, Sequence, Tuple

import torch
from torch import Tensor

from . import _lib


def wav_info(path: str) -> dict:
    sr, ch, fr, tag, bits = C.c_int(), C.c_int(), C.c_longlong(), C.c_int(), C.c_int()
    _lib.check(_lib.lib().rfx_wav_info(os.fsencode(path), C.byref(sr), C.byref(ch), C.byref(fr), C.byref(tag), C.byref(bits)), "rfx_wav_info")
    return {"sample_rate": sr.value, "channels": ch.value, "frames": fr.value, "format_tag": tag.value, "bits": bits.value}


def read_wav_batch(paths: Sequence[str], T: int, out: Tensor = None, threads: int = 8, pin: bool = True) -> Tuple[Tensor, List[int], List[int]]:
    """Decode `paths` (mono WAV files) into rows of a float32 CPU tensor (len(paths), 1, T); returns (batch, frames, sample_rates)."""
    n = len(paths)
    if n == 0:
        raise ValueError("no files given")
    if out is None:
        out = torch.empty(n, 1, T, dtype=torch.float32, pin_memory=pin and torch.cuda.is_available())
    if out.is_cuda or out.dtype != torch.float32 or tuple(out.shape) != (n, 1, T) or not out.is_contiguous():
        raise ValueError(f"out must be a contiguous float32 CPU tensor of shape ({n}, 1, {T})")
    arr = (C.c_char_p * n)(*[os.fsencode(p) for p in paths])
    frames = (C.c_longlong * n)()
    srs = (C.c_int * n)()
    rc = _lib.lib().rfx_ingest_wav_batch(arr, n, out.data_ptr(), T, int(threads), frames, srs)
    _lib.check(rc, "rfx_ingest_wav_batch")
    return out, list(frames), list(srs)


def read_wav(path: str) -> Tuple[Tensor, int]:
    """`torchaudio.load(path)` for a mono file: ((1, T) float32, sample_rate)."""
    info = wav_info(path)
    x, _, srs = read_wav_batch([path], max(1, info["frames"]), pin=False, threads=1)
    return x[0, :, : info["frames"]], srs[0]


class EffectChunkReader:
    """`len()` / `[idx]` of `EffectDataset` on a rendered chunk directory (remfx/datasets.py:458-468)."""

    def __init__(self, proc_root: str):
        self.proc_root = str(proc_root)
        self.total_chunks = len([d for d in os.listdir(self.proc_root) if os.path.isdir(os.path.join(self.proc_root, d))])

    def __len__(self) -> int:
        return self.total_chunks

    def paths(self, idx: int) -> Tuple[str, str, str, str]:
        d = os.path.join(self.proc_root, str(idx))
        return (os.path.join(d, "input.wav"), os.path.join(d, "target.wav"), os.path.join(d, "dry_effects.pt"), os.path.join(d, "wet_effects.pt"))

    def __getitem__(self, idx: int):
        if not 0 <= idx < self.total_chunks:
            raise IndexError(idx)
        pin, ptg, pdry, pwet = self.paths(idx)
        x, _ = read_wav(pin)
        y, _ = read_wav(ptg)
        return x, y, torch.load(pdry), torch.load(pwet)


class BatchIngest:
    """Iterates a chunk directory in batches of pinned tensors: (input (B,1,T), target (B,1,T), dry (B,5), wet (B,5)).

    `n_buffers` staging pairs are rotated, so a batch stays valid while the next `n_buffers - 1` are produced -- enough for the
    device pipeline, which holds a host buffer until the step has left stage 0 (use at least depth + 1)."""

    def __init__(self, reader: EffectChunkReader, batch_size: int, chunk_size: int = 262144, threads: int = 8, n_buffers: int = 4,
                 drop_last: bool = True, sample_rate: int = None):
        self.reader, self.B, self.T, self.threads = reader, int(batch_size), int(chunk_size), int(threads)
        self.drop_last, self.sample_rate = drop_last, sample_rate
        pin = torch.cuda.is_available()
        self._bufs = [(torch.empty(self.B, 1, self.T, pin_memory=pin), torch.empty(self.B, 1, self.T, pin_memory=pin)) for _ in range(n_buffers)]

    def __len__(self) -> int:
        n = len(self.reader)
        return n // self.B if self.drop_last else (n + self.B - 1) // self.B

    def __iter__(self) -> Iterator[Tuple[Tensor, Tensor, Tensor, Tensor]]:
        n = len(self.reader)
        for bi, start in enumerate(range(0, n, self.B)):
            idx = list(range(start, min(n, start + self.B)))
            if len(idx) < self.B and self.drop_last:
                return
            xb, yb = self._bufs[bi % len(self._bufs)]
            xb, yb = xb[: len(idx)], yb[: len(idx)]
            ps = [self.reader.paths(i) for i in idx]
            _, _, srx = read_wav_batch([p[0] for p in ps], self.T, out=xb, threads=self.threads)
            _, _, sry = read_wav_batch([p[1] for p in ps], self.T, out=yb, threads=self.threads)
            if self.sample_rate is not None and any(s != self.sample_rate for s in srx + sry):
                raise ValueError(f"chunk files are not all at {self.sample_rate} Hz")
            dry = torch.stack([torch.load(p[2]).float() for p in ps])
            wet = torch.stack([torch.load(p[3]).float() for p in ps])
            yield xb, yb, dry, wet

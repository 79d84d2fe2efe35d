"""Batch ingest: native WAV decode into pinned batches vs scipy / numpy, the EffectDataset item contract
(remfx/datasets.py:461-468) on a rendered chunk directory, and the error behaviour.  CPU only (the decoder is host code)."""
import os
import struct

import numpy as np
import pytest
import torch
from scipy.io import wavfile

from remfx_b200 import _lib
from remfx_b200.ingest import BatchIngest, EffectChunkReader, read_wav, read_wav_batch, wav_info


def _write_float_wav_sndfile_style(path, x, sr, extensible=False):
    """32-bit float WAV the way libsndfile (torchaudio.save's backend) lays it out: fmt, fact and PEAK chunks before data."""
    data = x.astype("<f4").tobytes()
    if extensible:
        guid = struct.pack("<H", 3) + bytes.fromhex("000000001000800000AA00389B71")
        fmt = struct.pack("<HHIIHHHHI", 0xFFFE, 1, sr, sr * 4, 4, 32, 22, 32, 4) + guid
    else:
        fmt = struct.pack("<HHIIHH", 3, 1, sr, sr * 4, 4, 32)
    fact = struct.pack("<I", len(x))
    peak = struct.pack("<IIfI", 1, 0, float(np.abs(x).max()), 0)
    odd = b"LIST" + struct.pack("<I", 5) + b"INFOx" + b"\0"   # odd-sized chunk with its pad byte
    body = b"WAVE" + b"fmt " + struct.pack("<I", len(fmt)) + fmt + b"fact" + struct.pack("<I", 4) + fact + \
        b"PEAK" + struct.pack("<I", len(peak)) + peak + odd + b"data" + struct.pack("<I", len(data)) + data
    with open(path, "wb") as fh:
        fh.write(b"RIFF" + struct.pack("<I", len(body)) + body)


def test_float32_wav_bit_exact(tmp_path):
    rng = np.random.default_rng(0)
    x = (0.1 * rng.standard_normal(5000)).astype(np.float32)
    for ext in (False, True):
        p = str(tmp_path / f"f{int(ext)}.wav")
        _write_float_wav_sndfile_style(p, x, 48000, extensible=ext)
        info = wav_info(p)
        assert info == {"sample_rate": 48000, "channels": 1, "frames": 5000, "format_tag": 3, "bits": 32}
        y, sr = read_wav(p)
        assert sr == 48000 and y.shape == (1, 5000) and y.dtype == torch.float32
        assert np.array_equal(y.numpy()[0], x)
    p = str(tmp_path / "scipy.wav")
    wavfile.write(p, 44100, x)
    sr, ref = wavfile.read(p)
    y, sr2 = read_wav(p)
    assert sr2 == sr == 44100 and np.array_equal(y.numpy()[0], ref)


@pytest.mark.parametrize("dtype,scale", [(np.int16, 32768.0), (np.int32, 2147483648.0), (np.uint8, None)])
def test_integer_pcm_matches_normalised_load(tmp_path, dtype, scale):
    rng = np.random.default_rng(1)
    if dtype == np.uint8:
        x = rng.integers(0, 256, 3000, dtype=np.uint8)
        ref = (x.astype(np.float32) - 128.0) / 128.0
    else:
        ii = np.iinfo(dtype)
        x = rng.integers(ii.min, ii.max, 3000, dtype=dtype)
        ref = (x.astype(np.float64) / scale).astype(np.float32)
    p = str(tmp_path / "i.wav")
    wavfile.write(p, 48000, x)
    y, _ = read_wav(p)
    assert np.array_equal(y.numpy()[0], ref)


def test_batch_rows_padding_and_threads(tmp_path):
    rng = np.random.default_rng(2)
    T = 4096
    sigs, paths = [], []
    for i in range(9):
        n = T if i != 4 else T - 100          # one short file: zero-padded row
        n = n if i != 7 else T + 50           # one long file: cut
        x = (0.1 * rng.standard_normal(n)).astype(np.float32)
        p = str(tmp_path / f"{i}.wav")
        _write_float_wav_sndfile_style(p, x, 48000)
        sigs.append(x)
        paths.append(p)
    for threads in (1, 4, 32):
        out, frames, srs = read_wav_batch(paths, T, threads=threads, pin=False)
        assert out.shape == (9, 1, T) and srs == [48000] * 9
        assert frames[4] == T - 100 and frames[7] == T + 50
        for i, x in enumerate(sigs):
            row = out[i, 0].numpy()
            m = min(len(x), T)
            assert np.array_equal(row[:m], x[:m]) and not row[m:].any()


def test_errors(tmp_path):
    with pytest.raises(ValueError, match="cannot open"):
        read_wav_batch([str(tmp_path / "missing.wav")], 16, pin=False)
    p = str(tmp_path / "junk.wav")
    open(p, "wb").write(b"not a wav file at all")
    with pytest.raises(ValueError, match="RIFF"):
        wav_info(p)
    st = str(tmp_path / "stereo.wav")
    wavfile.write(st, 48000, np.zeros((100, 2), np.float32))
    with pytest.raises(ValueError, match="mono"):
        read_wav_batch([st], 100, pin=False)
    tr = str(tmp_path / "trunc.wav")
    _write_float_wav_sndfile_style(tr, np.ones(1000, np.float32), 48000)
    blob = open(tr, "rb").read()
    open(tr, "wb").write(blob[:-400])
    with pytest.raises(ValueError, match="truncated"):
        read_wav_batch([tr], 1000, pin=False)
    with pytest.raises(ValueError):
        read_wav_batch([], 10)


def _render_dir(root, n, T, seed=3):
    """A chunk directory as remfx/datasets.py:197-200 writes it (float WAVs + 5-float label tensors)."""
    rng = np.random.default_rng(seed)
    items = []
    for i in range(n):
        d = root / str(i)
        d.mkdir()
        wet = (0.1 * rng.standard_normal(T)).astype(np.float32)
        dry = (0.1 * rng.standard_normal(T)).astype(np.float32)
        _write_float_wav_sndfile_style(str(d / "input.wav"), wet, 48000)
        _write_float_wav_sndfile_style(str(d / "target.wav"), dry, 48000)
        dl = torch.tensor(rng.integers(0, 2, 5), dtype=torch.float32)
        wl = torch.tensor(rng.integers(0, 2, 5), dtype=torch.float32)
        torch.save(dl, d / "dry_effects.pt")
        torch.save(wl, d / "wet_effects.pt")
        items.append((wet, dry, dl, wl))
    return items


def test_effect_chunk_reader_and_batches(tmp_path):
    T = 2048
    items = _render_dir(tmp_path, 7, T)
    rd = EffectChunkReader(str(tmp_path))
    assert len(rd) == 7
    x, y, dl, wl = rd[3]   # the EffectDataset.__getitem__ tuple
    assert x.shape == (1, T) and np.array_equal(x.numpy()[0], items[3][0]) and np.array_equal(y.numpy()[0], items[3][1])
    assert torch.equal(dl, items[3][2]) and torch.equal(wl, items[3][3])
    with pytest.raises(IndexError):
        rd[7]
    ing = BatchIngest(rd, batch_size=3, chunk_size=T, threads=2, n_buffers=2, sample_rate=48000)
    assert len(ing) == 2
    seen = 0
    for bi, (xb, yb, dry, wet) in enumerate(ing):
        assert xb.shape == (3, 1, T) and dry.shape == (3, 5)
        for j in range(3):
            it = items[3 * bi + j]
            assert np.array_equal(xb[j, 0].numpy(), it[0]) and np.array_equal(yb[j, 0].numpy(), it[1])
            assert torch.equal(dry[j], it[2]) and torch.equal(wet[j], it[3])
        seen += 3
    assert seen == 6   # drop_last
    assert len(BatchIngest(rd, 3, T, drop_last=False)) == 3

"""Shared helpers for the parity tests."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def relrms(a: torch.Tensor, b: torch.Tensor) -> float:
    """||a - b||_2 / ||b||_2 -- the parity metric of BASELINE.json (gate 1e-4, fp32)."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def golden(name: str):
    return np.load(os.path.join(GOLDEN, name))


def state_to(sd, device):
    return {k: v.to(device) for k, v in sd.items()}


def tcn_backward_case(g):
    """Inputs of tests/golden/tcn_backward.npz regenerated from its seeds (oracle/make_golden.py:tcn_backward_inputs):
    kink-free TCN weights (PReLU slopes = 1), audio, objective weights r; plus the reference's gradients keyed like its state_dict."""
    from oracle import weights

    nblocks, width, B, T = int(g["nblocks"]), int(g["width"]), int(g["B"]), int(g["T"])
    sd = weights.tcn_state(int(g["wseed"]), nblocks=nblocks, width=width)
    for k in sd:
        if k.endswith("relu.weight"):
            sd[k] = torch.ones_like(sd[k])
    assert abs(weights.checksum(sd) - float(g["wsum"])) < 1e-6 * abs(float(g["wsum"]))
    x = weights.synth_audio(int(g["xseed"]), B, T)
    out_ref = torch.from_numpy(g["out"])
    r = torch.randn(B, 1, out_ref.shape[-1], generator=torch.Generator().manual_seed(int(g["rseed"])))
    grads = {k[len("grad/"):]: torch.from_numpy(g[k]) for k in g.files if k.startswith("grad/")}
    return sd, x, r, out_ref, grads, dict(nblocks=nblocks, channel_width=width)


def example_case(g):
    """tests/golden/example_wav.npz: the reference repository's example.wav at 16 bits -> (1, 1, 262144) fp32 audio, plus the
    decimation step the stored outputs use (oracle/make_golden.py:example_wav_golden)."""
    x = (torch.from_numpy(g["pcm16"].astype("float32")) / 32767.0).reshape(1, 1, -1)
    return x, int(g["decim"])

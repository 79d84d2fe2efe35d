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

"""Recurrence kernel timing: B items x F steps, H=256, by slots per cluster."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from remfx_b200 import ops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
F = 513
H = 256
G = torch.randn(B * F, 8 * H, device="cuda") * 0.5
Whh = (torch.rand(2, 4 * H, H, device="cuda") * 2 - 1) * H ** -0.5
for impl, slots in (("mma", 8), ("tc", 16), ("tc", 32)):
    for _ in range(2):
        ops.lstm_layer(G, Whh, B, F, impl=impl, slots=slots)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.lstm_layer(G, Whh, B, F, impl=impl, slots=slots)
    e1.record()
    torch.cuda.synchronize()
    print(f"B={B} impl={impl} slots={slots} NACC={os.environ.get('RFX_LSTM_TC_NACC')}: {e0.elapsed_time(e1) / 10:.4f} ms per layer launch")

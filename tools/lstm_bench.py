"""Time one BiLSTM layer recurrence (B=32, F=513, H=256) in isolation (development aid)."""
import sys

import torch

sys.path.insert(0, ".")
from remfx_b200 import ops  # noqa: E402

import ctypes as C
from remfx_b200 import _lib
B, F, H = int(sys.argv[1]) if len(sys.argv) > 1 else 32, 513, 256
g = torch.Generator().manual_seed(0)
G = (torch.randn(B * F, 8 * H, generator=g) * 0.5).cuda()
Whh = ((torch.rand(2, 4 * H, H, generator=g) * 2 - 1) / 16).cuda()
mc, nb = C.c_int(), C.c_int()
_lib.lib().rfx_lstm_info(B, C.byref(mc), C.byref(nb))
print(f"max active clusters = {mc.value}, batch per cluster = {nb.value}")
for impl in ("mma", "ffma"):
    for _ in range(3):
        out = ops.lstm_layer(G, Whh, B, F, impl=impl)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 10
    e0.record()
    for _ in range(n):
        out = ops.lstm_layer(G, Whh, B, F, impl=impl)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"lstm layer [{impl}] B={B} F={F}: {ms:.3f} ms  ({ms * 1e3 / F:.3f} us/step, {ms * 1e-3 / F * 1.965e9:.0f} cycles/step @1.965GHz)  checksum {float(out.sum()):.4f}")

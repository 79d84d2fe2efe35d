"""Time the tcgen05 dense engine on the Open-Unmix layer shapes (development aid)."""
import sys

import torch

sys.path.insert(0, ".")
from remfx_b200 import ops  # noqa: E402

shapes = {"wih": (16416, 2048, 512), "fc2": (16416, 512, 1024), "fc3": (16416, 1025, 512), "fc1": (16416, 512, 1025)}
which = sys.argv[1].split(",") if len(sys.argv) > 1 else list(shapes)
g = torch.Generator().manual_seed(0)
for name in which:
    M, N, K = shapes[name]
    A = torch.randn(M, K, generator=g).cuda()
    W = (torch.randn(N, K, generator=g) / K ** 0.5).cuda()
    t1 = torch.randn(N, generator=g).cuda()
    for _ in range(2):
        out = ops.linear(A, W, t1=t1, impl="tc")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 5
    e0.record()
    for _ in range(n):
        out = ops.linear(A, W, t1=t1, impl="tc")
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"{name}: M={M} N={N} K={K}: {ms:.3f} ms/call incl. operand split ({2 * M * N * K / ms / 1e9:.1f} TFLOP/s fp32-equivalent)")

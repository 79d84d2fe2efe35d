"""Where does the host-buffer pipeline lose time?  Host time per push, PCIe copy rates, e2e ms/step by consumer lag."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from remfx_b200.models import OpenUnmixModel  # noqa: E402
from remfx_b200.synth import synth_audio  # noqa: E402

B, T, K = 32, 262144, 40
torch.manual_seed(0)
m = OpenUnmixModel(n_fft=2048, hop_length=512, n_channels=1, alpha=0.3, sample_rate=48000).cuda().eval()
xh = [synth_audio(10 + i, B, T).pin_memory() for i in range(5)]
oh = [torch.empty(B, 1, T).pin_memory() for _ in range(8)]
xd = [x.cuda() for x in xh]
od = [torch.empty_like(xd[0]) for _ in range(8)]
# raw PCIe
for name, src, dst in (("H2D", xh[0], xd[0]), ("D2H", xd[0], oh[0])):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 10
    print(f"{name}: {src.numel() * 4 / dt / 1e9:.1f} GB/s ({dt * 1e3:.3f} ms per 33.5 MB)")
pipe = m.pipeline()
for mode, xs, outs in (("device", xd, od), ("host", xh, oh)):
    for lag in (3, 4, 6):
        for i in range(4):
            pipe.push(xs[i], outs[i])
        pipe.flush()
        torch.cuda.synchronize()
        seqs, push_t = [], []
        t0 = time.perf_counter()
        for k in range(K):
            if k >= lag:
                pipe.wait(seqs[k - lag])
            a = time.perf_counter()
            seqs.append(pipe.push(xs[k % 5], outs[k % 8]))
            push_t.append(time.perf_counter() - a)
        pipe.flush()
        for sq in seqs[-lag:]:
            pipe.wait(sq)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / K
        push_t.sort()
        print(f"{mode:6s} lag={lag}: {dt * 1e3:.3f} ms/step; host time per push: median {push_t[len(push_t) // 2] * 1e3:.3f} ms, max {push_t[-1] * 1e3:.3f} ms")

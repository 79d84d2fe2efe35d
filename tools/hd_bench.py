"""Time the Hybrid-Demucs forward on 262144-sample chunks (development aid)."""
import sys

import torch

sys.path.insert(0, ".")
from remfx_b200.models import DemucsModel  # noqa: E402
from remfx_b200.synth import synth_audio  # noqa: E402

T = 262144
torch.manual_seed(0)
m = DemucsModel(sample_rate=48000, sources=["mixture"], audio_channels=1, nfft=4096, channels=48).cuda().eval()
for B in [int(a) for a in sys.argv[1:]] or [1, 8]:
    x = synth_audio(1, B, T).cuda()
    for _ in range(2):
        y = m.sample(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 5
    e0.record()
    for _ in range(n):
        y = m.sample(x)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"HDemucs B={B}: {ms:.2f} ms/call, {B * T / 48000 / (ms / 1e3):.1f} audio-s/s, {117.0 * B / ms:.1f} TFLOP/s fp32-equivalent, "
          f"{m.launches_per_call(B, T)} launches, workspace {m._ws.numel() / 2**30:.2f} GiB")

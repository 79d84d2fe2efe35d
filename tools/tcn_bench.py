"""Time the TCN forward (20 blocks, C=256) on 262144-sample chunks (development aid)."""
import sys

import torch

sys.path.insert(0, ".")
from remfx_b200.models import TCNModel  # noqa: E402
from remfx_b200.synth import synth_audio  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
T = 262144
torch.manual_seed(0)
m = TCNModel(sample_rate=48000, num_bins=1025, ninputs=1, noutputs=1, nblocks=20, channel_growth=0, channel_width=256, kernel_size=7,
             stack_size=10, dilation_growth=2, causal=False).cuda().eval()
x = synth_audio(1, B, T).cuda()
for _ in range(2):
    y = m.sample(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 3
e0.record()
for _ in range(n):
    y = m.sample(x)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
fl = 5135.5e9 * B
print(f"TCN B={B}: {ms:.2f} ms/call, {B * T / 48000 / (ms / 1e3):.1f} audio-s/s, {fl / ms / 1e9:.1f} TFLOP/s fp32-equivalent ({3 * fl / ms / 1e9:.0f} bf16-equivalent), out {tuple(y.shape)}")

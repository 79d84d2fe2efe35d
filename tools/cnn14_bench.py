"""Time the Cnn14 effect classifier (remfx/classifier.py:193-233) on 262144-sample chunks (development aid)."""
import sys

import torch

sys.path.insert(0, ".")
from remfx_b200.classifier import Cnn14  # noqa: E402
from remfx_b200.synth import synth_audio  # noqa: E402

T = 262144
torch.manual_seed(0)
m = Cnn14(num_classes=5, sample_rate=48000, model_sample_rate=48000, n_fft=2048, hop_length=512, n_mels=128).cuda().eval()
# forward FLOPs per chunk (fp32-equivalent): 3x3 convs over the pooled mel grid + fc1
F, Mel = T // 512 + 1, 128
chans = [1, 64, 128, 256, 512, 1024, 2048]
fl, h, w = 0.0, F, Mel
for i in range(6):
    fl += 2.0 * 9 * h * w * (chans[i] * chans[i + 1] + chans[i + 1] * chans[i + 1])
    if i < 5:
        h, w = h // 2, w // 2
fl += 2.0 * 2048 * 2048
for B in [int(a) for a in sys.argv[1:]] or [1, 16, 64]:
    x = synth_audio(1, B, T).cuda()
    for _ in range(2):
        y = m(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 5
    e0.record()
    for _ in range(n):
        y = m(x)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"Cnn14 B={B}: {ms:.2f} ms/call, {B * T / 48000 / (ms / 1e3):.1f} audio-s/s, {B * fl / ms / 1e9:.1f} TFLOP/s fp32-equivalent "
          f"({3 * B * fl / ms / 1e9:.0f} bf16-equivalent)")

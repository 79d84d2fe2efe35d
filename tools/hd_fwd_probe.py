"""One Hybrid-Demucs forward (+ loss) at B x 262144 for an ncu launch list: python tools/hd_fwd_probe.py [B]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from remfx_b200.models import DemucsModel  # noqa: E402
from remfx_b200.synth import synth_audio  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
torch.manual_seed(0)
m = DemucsModel(sample_rate=48000, sources=["mixture"], audio_channels=1, nfft=4096, channels=48).cuda().eval()
x, y = synth_audio(1, B, 262144).cuda(), synth_audio(2, B, 262144).cuda()
with torch.no_grad():
    for _ in range(2):
        loss, out = m((x, y))
torch.cuda.synchronize()
print(float(loss))

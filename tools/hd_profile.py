"""One profiled HDemucs forward (ncu --profile-from-start off ...)."""
import sys

import torch

sys.path.insert(0, ".")
from remfx_b200.models import DemucsModel  # noqa: E402
from remfx_b200.synth import synth_audio  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
torch.manual_seed(0)
m = DemucsModel(sample_rate=48000, sources=["mixture"], audio_channels=1, nfft=4096, channels=48).cuda().eval()
x = synth_audio(1, B, 262144).cuda()
m.sample(x)
torch.cuda.synchronize()
torch.cuda.profiler.start()
m.sample(x)
torch.cuda.synchronize()
torch.cuda.profiler.stop()

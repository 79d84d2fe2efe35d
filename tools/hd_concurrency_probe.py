"""How much of a Hybrid-Demucs forward leaves the GPU under-filled?  Two independent forwards (two model instances, two streams)
against the same two forwards back to back: the ratio bounds what overlapping the frequency and time branches could give."""
import sys

import torch

sys.path.insert(0, ".")
from remfx_b200.models import DemucsModel  # noqa: E402
from remfx_b200.synth import synth_audio  # noqa: E402

T = 262144
for B in [int(a) for a in sys.argv[1:]] or [1, 16]:
    torch.manual_seed(0)
    ms_ = [DemucsModel(sample_rate=48000, sources=["mixture"], audio_channels=1, nfft=4096, channels=48).cuda().eval() for _ in range(2)]
    xs = [synth_audio(1 + i, B, T).cuda() for i in range(2)]
    st = [torch.cuda.Stream() for _ in range(2)]
    for m, x in zip(ms_, xs):
        for _ in range(2):
            m.sample(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 5
    e0.record()
    for _ in range(n):
        for m, x in zip(ms_, xs):
            m.sample(x)
    e1.record()
    torch.cuda.synchronize()
    seq = e0.elapsed_time(e1) / n
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        for m, x, s in zip(ms_, xs, st):
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                m.sample(x)
        for s in st:
            torch.cuda.current_stream().wait_stream(s)
    e1.record()
    torch.cuda.synchronize()
    con = e0.elapsed_time(e1) / n
    print(f"B={B}: two forwards back to back {seq:.2f} ms, on two streams {con:.2f} ms ({seq / con:.2f}x)")

"""Quick device-side timing of the Open-Unmix path with per-stage breakdown (development aid; bench.py is the contract)."""
import sys

import torch

sys.path.insert(0, ".")
from remfx_b200.models import OpenUnmixModel  # noqa: E402
from remfx_b200.synth import synth_audio  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
impls = sys.argv[2].split(",") if len(sys.argv) > 2 else ["tc"]
T = 262144
for impl in impls:
    torch.manual_seed(0)
    m = OpenUnmixModel(sample_rate=48000).cuda().eval()
    x = synth_audio(1, B, T).cuda()
    m.set_profiling(True)
    for _ in range(3):
        m.sample(x)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    n = 10
    ev[0].record()
    for _ in range(n):
        m.sample(x)
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / n
    st = m.stage_times_ms()
    print(f"impl={impl} B={B} ms/call={ms:.3f} audio_s_per_s={B * T / 48000 / (ms / 1e3):.1f}", flush=True)
    print("  " + " ".join(f"{k}={v:.3f}" for k, v in st.items()), flush=True)

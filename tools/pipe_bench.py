"""Pipeline vs serial timing of OpenUnmixModel at B=32 x 262144 (device-resident).  Env: RFX_UMX_PIPE_MAX_SMS / RFX_UMX_PIPE_SLOTS."""
import os
import sys
import statistics

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from remfx_b200.models import OpenUnmixModel  # noqa: E402
from remfx_b200.synth import synth_audio  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
K = int(sys.argv[2]) if len(sys.argv) > 2 else 40
T = 262144
torch.manual_seed(0)
if os.environ.get("RFX_LSTM_IMPL"):
    from remfx_b200 import _lib
    _lib.check(_lib.lib().rfx_lstm_set_impl(int(os.environ["RFX_LSTM_IMPL"])), "set_impl")
m = OpenUnmixModel(n_fft=2048, hop_length=512, n_channels=1, alpha=0.3, sample_rate=48000).cuda().eval()
xs = [synth_audio(10 + i, B, T).cuda() for i in range(5)]
outs = [torch.empty_like(xs[0]) for _ in range(6)]
for i in range(3):
    m.sample(xs[i])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for k in range(K):
    m.sample(xs[k % 5])
e1.record()
torch.cuda.synchronize()
print(f"serial  : {e0.elapsed_time(e1) / K:.3f} ms/step")
pipe = m.pipeline()
for i in range(4):
    pipe.push(xs[i], outs[i])
pipe.flush()
torch.cuda.synchronize()
pipe.set_profiling(3 * K)
e0.record()
for k in range(K):
    pipe.push(xs[k % 5], outs[k % 6])
pipe.flush()
e1.record()
torch.cuda.synchronize()
rt = pipe.recurrence_times_ms()
print(f"pipeline: {e0.elapsed_time(e1) / K:.3f} ms/step  (recurrence launches: n={len(rt)} mean {statistics.mean(rt):.3f} median {statistics.median(rt):.3f} "
      f"max {max(rt):.3f} ms) {pipe.info()}  env MAX_SMS={os.environ.get('RFX_UMX_PIPE_MAX_SMS')} SLOTS={os.environ.get('RFX_UMX_PIPE_SLOTS')} IMPL={os.environ.get('RFX_LSTM_IMPL')} STREAMS={os.environ.get('RFX_UMX_PIPE_REC_STREAMS')}")

#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_loss.py -m gpu -q --timeout 200 --no-header -p no:cacheprovider > gpurun_out/t.log 2>&1; echo "loss tests exit=$? $(tail -n 1 gpurun_out/t.log)"; grep -E "^FAILED|^ERROR|rror|assert" gpurun_out/t.log | head -12
python - <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd())
from remfx_b200.losses import remfx_loss
from remfx_b200.synth import synth_audio
B, T = 32, 262144
a = synth_audio(1, B, T).cuda().requires_grad_(True); b = synth_audio(2, B, T).cuda()
for _ in range(2):
    l = remfx_loss(a, b); l.backward()
torch.cuda.synchronize()
e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
e[0].record(); l = remfx_loss(a, b); e[1].record(); l.backward(); e[2].record(); torch.cuda.synchronize()
print(f"loss fwd {e[0].elapsed_time(e[1]):.3f} ms, bwd {e[1].elapsed_time(e[2]):.3f} ms at {B}x{T}")
PY

#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_gemm_lstm.py -m gpu -q --timeout 120 --no-header -p no:cacheprovider -k lstm > gpurun_out/test_lstm.log 2>&1; echo "tests exit=$? $(tail -n 1 gpurun_out/test_lstm.log)"
grep -E "FAILED|Error|error|timed out" gpurun_out/test_lstm.log | head -20
timeout 120 python tools/lstm_bench.py 32 2>&1 | tail -5
for cfg in "16 1" "16 2" "24 2" "32 2" "32 3"; do set -- $cfg
RFX_UMX_PIPE_SLOTS=$1 RFX_UMX_PIPE_REC_STREAMS=$2 timeout 120 python tools/pipe_bench.py 32 40 2>&1 | tail -1 | sed "s/^/slots=$1 streams=$2 /"
done
cat > /tmp/one.py <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd())
from remfx_b200 import ops
B, F, H = 32, 513, 256
G = torch.randn(B * F, 8 * H, device="cuda") * 0.5
Whh = (torch.rand(2, 4 * H, H, device="cuda") * 2 - 1) * H ** -0.5
for _ in range(3):
    ops.lstm_layer(G, Whh, B, F, slots=16)
torch.cuda.synchronize()
PY
ncu --set full --clock-control none --import-source on -k regex:lstm_rec_ws -s 2 -c 1 -o gpurun_out/prof_lstm_ws python /tmp/one.py > gpurun_out/ncu_ws.log 2>&1; echo "ncu exit=$?"

#!/bin/bash
mkdir -p gpurun_out
python bench.py --warmup 3 > gpurun_out/bench_k200.json 2> gpurun_out/bench_k200.err; echo "bench K=200 exit=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_k200.json')); print('value',d['value'], d['ms_per_step'], 'steps',d['steps'],'e2e', d['e2e']['value'], d['e2e']['ms_per_step'])"
cat > /tmp/one.py <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd())
from remfx_b200 import ops
B, F, H = 32, 513, 256
G = torch.randn(B * F, 8 * H, device="cuda") * 0.5
Whh = (torch.rand(2, 4 * H, H, device="cuda") * 2 - 1) * H ** -0.5
for _ in range(3):
    ops.lstm_layer(G, Whh, B, F, impl="tc", slots=32)
torch.cuda.synchronize()
PY
ncu --set full --clock-control none --import-source on -k regex:lstm_rec_tc -s 2 -c 1 -o gpurun_out/prof_lstm_tc32 python /tmp/one.py > gpurun_out/ncu_tc32.log 2>&1; echo "ncu exit=$?"

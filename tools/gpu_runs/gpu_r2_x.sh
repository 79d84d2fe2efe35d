#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2x_bench_n2.json 2> gpurun_out/r2x_bench_n2.err; echo "bench n2 exit=$?"
python - <<'P'
import json
try:
    d=json.loads(open("gpurun_out/r2x_bench_n2.json").read().strip().splitlines()[-1])
    print("n", d["n_gpus"], "value", round(d["value"]), "ms", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]))
    for o in d.get("other_configs",[]): print(o.get("config","")[:60], o.get("ms_per_step"), o.get("value"), o.get("all_reduce_ms"))
except Exception as e:
    print("ERR", e); print(open("gpurun_out/r2x_bench_n2.err").read()[-2000:])
P
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/umx_train_bench.py --batch 16 --steps 3 --warmup 2 2>/dev/null | cut -c1-500

#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2y_bench_n8.json 2> gpurun_out/r2y_bench_n8.err; echo "bench n8 exit=$?"
python - <<'P'
import json
try:
    d=json.loads(open("gpurun_out/r2y_bench_n8.json").read().strip().splitlines()[-1])
    print("n", d["n_gpus"], "value", round(d["value"]), "ms", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]), d["e2e"].get("ms_per_step"))
    for o in d.get("other_configs",[]): print(o.get("config","")[:60], o.get("ms_per_step"), o.get("value"), {k:v for k,v in o.items() if "reduce" in k})
except Exception as e:
    print("ERR", e); print(open("gpurun_out/r2y_bench_n8.err").read()[-2000:])
P

#!/bin/bash
# end-of-round verification: full GPU suite, then the 2-GPU bench (training leg with the real all-reduce)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2f_tests.log 2>&1; echo "tests exit=$?"; tail -2 gpurun_out/r2f_tests.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2f_bench_n2.json 2> gpurun_out/r2f_bench_n2.err; echo "bench n2 exit=$?"
python - <<'P'
import json
d=json.loads(open("gpurun_out/r2f_bench_n2.json").read().strip().splitlines()[-1])
print("n", d["n_gpus"], "value", round(d["value"]), "ms", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]))
for o in d.get("other_configs",[]): print(o.get("config","")[:50], o.get("ms_per_step"), o.get("value"), {k:(round(v,3) if isinstance(v,float) else v) for k,v in (o.get("all_reduce") or {}).items() if k in ("ms","ms_last_rank_to_arrive","busbw_GBs")})
P

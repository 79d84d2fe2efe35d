#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r3z_smoke.log 2>&1; echo "smoke exit=$?"; tail -3 gpurun_out/r3z_smoke.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r3z_bench_n4.json 2> gpurun_out/r3z_bench_n4.err; echo "bench n4 exit=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29552 bench.py --impl reference --gpus 4 --steps 2 --warmup 1 > gpurun_out/r3z_bench_ref_n4.json 2> gpurun_out/r3z_bench_ref_n4.err; echo "ref arm n4 exit=$?"; cut -c1-200 gpurun_out/r3z_bench_ref_n4.json
python - <<'P'
import json
d=json.loads(open("gpurun_out/r3z_bench_n4.json").read().strip().splitlines()[-1])
print("n", d["n_gpus"], "value", round(d["value"]), "ms", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]))
for o in d.get("other_configs",[]): print(o.get("config","")[:50], o.get("ms_per_step"), o.get("value"), {k:(round(v,3) if isinstance(v,float) else v) for k,v in (o.get("all_reduce") or {}).items() if k in ("ms","ms_last_rank_to_arrive","busbw_GBs")})
P

#!/bin/bash
# session-3 verification at HEAD: full GPU suite, then the driver's bench line (N=1) and the reference arm
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r3a_tests.log 2>&1; echo "tests exit=$?"; tail -3 gpurun_out/r3a_tests.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r3a_bench_k20.json 2> gpurun_out/r3a_bench_k20.err; echo "bench exit=$?"
python - <<'P'
import json
d=json.loads(open("gpurun_out/r3a_bench_k20.json").read().strip().splitlines()[-1])
print("value", round(d["value"]), "ms", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]), "rec ms", round(d["roofline"]["ms_per_launch"],4))
for o in d.get("other_configs",[]): print(o.get("config","")[:60], o.get("ms_per_step"), o.get("value"))
P

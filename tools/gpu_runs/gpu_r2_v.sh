#!/bin/bash
# verification after the staged STFT / locked recurrence / Open-Unmix training / tile-width changes
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2v_tests.log 2>&1; echo "tests exit=$?"; tail -3 gpurun_out/r2v_tests.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2v_bench_k20.json 2> gpurun_out/r2v_bench_k20.err; echo "bench exit=$?"
timeout 600 python tools/umx_train_bench.py --batch 16 --steps 4 --warmup 2 > gpurun_out/r2v_umx_train_b16.json 2> gpurun_out/r2v_umx.err; echo "umx train exit=$?"
python - <<'P'
import json
d=json.loads(open("gpurun_out/r2v_bench_k20.json").read().strip().splitlines()[-1])
print("value", round(d["value"]), "ms", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]), "rec ms", round(d["roofline"]["ms_per_launch"],4))
for o in d.get("other_configs",[]): print(o.get("config","")[:60], o.get("ms_per_step"), o.get("value"))
for f in d["roofline"].get("kernel_families_serial",[]): print(f["kernel"], f["ms"], round(f["achieved"],1), f["unit"], round(f["frac"],3))
print(open("gpurun_out/r2v_umx_train_b16.json").read()[:400])
P

#!/bin/bash
# full GPU suite + bench (K = 20 as the driver runs it, K = 200) + ncu captures of the reworked kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2q_tests.log 2>&1; echo "tests exit=$?"; tail -3 gpurun_out/r2q_tests.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2q_bench_k20.json 2> gpurun_out/r2q_bench_k20.err; echo "bench exit=$?"
timeout 600 python bench.py --steps 200 --warmup 5 --legs none --no-cpu-baseline --no-gpu-eager > gpurun_out/r2q_bench_k200.json 2> gpurun_out/r2q_bench_k200.err; echo "bench200 exit=$?"
N="ncu --set full --clock-control none --import-source on -f"
timeout 300 $N -k regex:"stft2048_tma_kernel|lstm_rec_tc_kernel" -s 6 -c 3 -o gpurun_out/r2q_umx python bench.py --steps 4 --warmup 3 --legs none --no-cpu-baseline --no-gpu-eager > gpurun_out/r2q_ncu.log 2>&1; echo "ncu exit=$?"
python - <<'P'
import json
for f in ("gpurun_out/r2q_bench_k20.json","gpurun_out/r2q_bench_k200.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value", round(d["value"]), "ms", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]), "rec ms", round(d["roofline"]["ms_per_launch"],4), "fast", d.get("bf16_fast",{}).get("ms_per_step"))
    except Exception as e: print(f, "ERR", e)
P

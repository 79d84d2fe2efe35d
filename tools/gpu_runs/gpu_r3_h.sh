#!/bin/bash
# gemm2: idle epilogue warps stay out of the tile loop, drain wait backs off
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_hdemucs.py tests/test_gpu_gemm_lstm.py tests/test_gpu_umx.py tests/test_gpu_tcn.py tests/test_gpu_cnn14.py -x -q > gpurun_out/r3h_tests.log 2>&1; echo "tests exit=$?"; tail -3 gpurun_out/r3h_tests.log
timeout 300 python tools/hd_bench.py 1 16 32 2>&1 | grep HDemucs | tee gpurun_out/r3h_hd_fwd.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r3h_bench_k20.json 2> gpurun_out/r3h_bench_k20.err; echo "bench exit=$?"
python - <<'P'
import json
d=json.loads(open("gpurun_out/r3h_bench_k20.json").read().strip().splitlines()[-1])
print("value", round(d["value"]), "ms", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]), "rec ms", round(d["roofline"]["ms_per_launch"],4))
for o in d.get("other_configs",[]): print(o.get("config","")[:60], o.get("ms_per_step"), o.get("value"))
for f in d["roofline"].get("kernel_families_serial",[]): print(f["kernel"], f["ms"], round(f["frac"],3))
P

#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm_lstm.py tests/test_gpu_umx.py -m gpu -q --timeout 120 --no-header -p no:cacheprovider > gpurun_out/r2i_tests.log 2>&1
echo "tests exit=$? $(tail -n 1 gpurun_out/r2i_tests.log)"
grep -E "^(FAILED|ERROR)|^E  " gpurun_out/r2i_tests.log | head -20
for mode in dual single; do
  for k in 20 200; do
    if [ $mode = single ]; then export RFX_LSTM_TC32_SINGLE=1; else unset RFX_LSTM_TC32_SINGLE; fi
    timeout 300 python bench.py --steps $k --warmup 5 --legs none --no-cpu-baseline --no-gpu-eager > gpurun_out/r2i_bench_${mode}_k$k.json 2> gpurun_out/r2i_bench_${mode}_k$k.err
    python - <<PY
import json
l = json.loads(open('gpurun_out/r2i_bench_${mode}_k$k.json').read().strip().splitlines()[-1])
print('$mode K=$k', 'ms/step', round(l['ms_per_step'],4), 'e2e ms', round(l['e2e']['ms_per_step'],4), 'rec ms/launch', round(l['roofline']['ms_per_launch'],4), 'us/step', round(l['roofline']['us_per_dependent_step'],3), 'bf16', l['bf16_fast'].get('ms_per_step'))
PY
  done
done

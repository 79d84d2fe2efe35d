#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 --no-header -p no:cacheprovider > gpurun_out/r2h_tests.log 2>&1
echo "tests exit=$? $(tail -n 1 gpurun_out/r2h_tests.log)"
grep -E "^(FAILED|ERROR)|^E  " gpurun_out/r2h_tests.log | head -20
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err
echo "bench exit=$?"; tail -n 3 gpurun_out/r2h_bench.err
python - <<'PY'
import json
l = json.loads(open('gpurun_out/r2h_bench.json').read().strip().splitlines()[-1])
print('value', l['value'], 'ms', l['ms_per_step'], 'e2e', l['e2e']['value'])
print('bf16_fast', l.get('bf16_fast'))
print('gpu_eager', {k: (v.get('ms_per_step') if isinstance(v, dict) else v) for k, v in l.get('gpu_eager', {}).items()})
for o in l.get('other_configs', []):
    print(o.get('config', '')[:60], o.get('ms_per_step'), o.get('value'), o.get('error'))
PY
timeout 600 python tools/hd_train_bench.py --batch 16 --steps 3 --warmup 2 > gpurun_out/r2h_hd_train_b16.json 2>/dev/null; cat gpurun_out/r2h_hd_train_b16.json

#!/bin/bash
# Run every GPU test file in its own process (a trapped kernel poisons the CUDA context) and keep the logs.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
for f in tests/test_gpu_*.py; do
  n=$(basename $f .py)
  timeout 600 python -m pytest $f -m gpu -q --timeout 300 -x --no-header -p no:cacheprovider > gpurun_out/$n.log 2>&1
  echo "$n exit=$?" | tee -a gpurun_out/summary.txt
  tail -n 25 gpurun_out/$n.log
done
for f in tests/test_gpu_*.py; do
  n=$(basename $f .py)
  timeout 900 python -m pytest $f -m gpu -q --timeout 300 --no-header -p no:cacheprovider > gpurun_out/${n}_all.log 2>&1
  echo "${n}_all exit=$?" | tee -a gpurun_out/summary.txt
  grep -E "passed|failed|FAILED|Error" gpurun_out/${n}_all.log | tail -n 30
done
timeout 600 python tools/umx_quick_bench.py 32 > gpurun_out/quick_bench.log 2>&1; cat gpurun_out/quick_bench.log | tail -5

#!/bin/bash
# Full round-end style validation: GPU tests (one process per file), smoke, bench (both arms), ncu launch list.
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
for f in tests/test_gpu_*.py; do
  n=$(basename $f .py)
  timeout 900 python -m pytest $f -m gpu -q --timeout 300 --no-header -p no:cacheprovider > gpurun_out/$n.log 2>&1
  echo "$n exit=$? $(tail -n 1 gpurun_out/$n.log)" | tee -a gpurun_out/summary.txt
done
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit=$? $(tail -n 1 gpurun_out/smoke.log)" | tee -a gpurun_out/summary.txt
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref exit=$?"
python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit=$?"; cat gpurun_out/bench.json; tail -n 3 gpurun_out/bench.err
timeout 900 python tools/bench_configs.py > gpurun_out/bench_configs.jsonl 2> gpurun_out/bench_configs.err; echo "bench_configs exit=$?"; cat gpurun_out/bench_configs.jsonl; tail -n 5 gpurun_out/bench_configs.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches exit=$?"

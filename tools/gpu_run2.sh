#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_loss.py -m gpu -q --timeout 300 --no-header -p no:cacheprovider > gpurun_out/test_gpu_loss.log 2>&1; echo "loss tests exit=$?"; tail -n 15 gpurun_out/test_gpu_loss.log
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit=$?"; tail -n 3 gpurun_out/smoke.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit=$?"; cat gpurun_out/bench.json; tail -n 5 gpurun_out/bench.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref exit=$?"; cat gpurun_out/bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches exit=$?"
ncu --set full --clock-control none --import-source on -k regex:lstm_rec -s 6 -c 1 -o gpurun_out/prof_lstm python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_lstm.log 2>&1; echo "ncu lstm exit=$?"
ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 10 -c 2 -o gpurun_out/prof_gemm python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_gemm.log 2>&1; echo "ncu gemm exit=$?"
ls -la gpurun_out

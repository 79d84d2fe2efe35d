#!/bin/bash
# Round-2 first GPU call: whole GPU suite (all failures reported), reference arm, bench with the new legs, tcgen05 wgrad probe.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 --no-header -p no:cacheprovider > gpurun_out/r2a_tests.log 2>&1
echo "tests exit=$? $(tail -n 1 gpurun_out/r2a_tests.log)"
grep -E "^(FAILED|ERROR)" gpurun_out/r2a_tests.log | head -20
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2a_bench_ref.json 2> gpurun_out/r2a_bench_ref.err
echo "ref exit=$?"; cut -c1-400 gpurun_out/r2a_bench_ref.json
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
echo "bench exit=$?"; tail -n 3 gpurun_out/r2a_bench.err; cut -c1-3000 gpurun_out/r2a_bench.json
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --expt-relaxed-constexpr -o /tmp/wgrad_tc_probe tools/wgrad_tc_probe.cu -lcuda \
  && timeout 120 /tmp/wgrad_tc_probe > gpurun_out/r2a_wgrad_tc_probe.log 2>&1; echo "probe exit=$?"; tail -n 15 gpurun_out/r2a_wgrad_tc_probe.log
nvidia-smi topo -m > gpurun_out/r2a_topo.txt 2>&1; nproc; lscpu | head -20 > gpurun_out/r2a_lscpu.txt

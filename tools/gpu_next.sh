#!/bin/bash
# First GPU call of the next session: everything written after the round-1 GPU budget ran out, in order of importance.
#   gpurun --timeout 900 -- 'bash tools/gpu_next.sh'
mkdir -p gpurun_out
# 1. the GPU tests that have not run on hardware yet (train-step harness, gradient golden), then the TCN files again
for f in tests/test_gpu_zz_example_wav.py tests/test_gpu_zz_train_step.py tests/test_gpu_tcn_backward.py tests/test_gpu_tcn.py; do
  n=$(basename $f .py)
  timeout 300 python -m pytest $f -m gpu -q --timeout 200 --no-header -p no:cacheprovider > gpurun_out/$n.log 2>&1
  echo "$n exit=$? $(tail -n 1 gpurun_out/$n.log)"
done
# 2. tcgen05 weight-gradient probe (MN-major SW128 operands): correctness vs the SIMT sums, then time at the benchmark length
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --expt-relaxed-constexpr -o /tmp/wgrad_tc_probe tools/wgrad_tc_probe.cu -lcuda \
  && timeout 120 /tmp/wgrad_tc_probe > gpurun_out/wgrad_tc_probe.log 2>&1; echo "probe exit=$?"; cat gpurun_out/wgrad_tc_probe.log
# 3. TCN training step with the CPU leg beside it (and the launch list / ncu capture)
bash tools/gpu_tcn_train.sh

#!/bin/bash
# HDemucs backward bring-up: forward regression, then the backward tests with blocking launches (errors land on the right label)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_hdemucs.py -m gpu -q --timeout 300 --no-header -p no:cacheprovider -x > gpurun_out/r2b_fwd.log 2>&1
echo "forward exit=$? $(tail -n 1 gpurun_out/r2b_fwd.log)"
CUDA_LAUNCH_BLOCKING=1 timeout 900 python -m pytest tests/test_gpu_hdemucs_backward.py -m gpu -q --timeout 600 --no-header -p no:cacheprovider -s > gpurun_out/r2b_bwd.log 2>&1
echo "backward exit=$? $(tail -n 1 gpurun_out/r2b_bwd.log)"
grep -E "^(FAILED|ERROR)|Error|error" gpurun_out/r2b_bwd.log | head -20

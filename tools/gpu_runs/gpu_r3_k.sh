#!/bin/bash
# fused-tap weight-gradient contraction for the narrow layers (swapped roles, all taps' accumulators in tensor memory)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_hdemucs_backward.py tests/test_gpu_umx_train.py -x -q > gpurun_out/r3k_tests.log 2>&1; echo "tests exit=$?"; tail -5 gpurun_out/r3k_tests.log
timeout 600 python tools/hd_train_bench.py --batch 16 --steps 4 --warmup 2 > gpurun_out/r3k_hd_train.json 2> gpurun_out/r3k_hd.err; echo "hd train exit=$?"; cut -c1-420 gpurun_out/r3k_hd_train.json
RFX_HD_WGRAD_FUSED=0 timeout 600 python tools/hd_train_bench.py --batch 16 --steps 4 --warmup 2 > gpurun_out/r3k_hd_train_unfused.json 2> gpurun_out/r3k_hd2.err; echo "hd train (unfused) exit=$?"; cut -c1-420 gpurun_out/r3k_hd_train_unfused.json

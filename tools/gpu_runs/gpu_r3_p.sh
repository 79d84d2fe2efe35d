#!/bin/bash
# gemm2 tap groups: one A box of 136 pixels serves the taps that differ by an x offset (descriptor base offset)
mkdir -p gpurun_out
RFX_G2_TAPGROUPS=1 timeout 900 python -m pytest tests/test_gpu_hdemucs.py tests/test_gpu_hdemucs_backward.py tests/test_gpu_gemm_lstm.py tests/test_gpu_tcn.py tests/test_gpu_tcn_backward.py tests/test_gpu_cnn14.py -x -q > gpurun_out/r3p_tests.log 2>&1; echo "tests exit=$?"; tail -5 gpurun_out/r3p_tests.log
RFX_G2_TAPGROUPS=1 timeout 300 python tools/hd_bench.py 1 16 32 2>&1 | grep HDemucs | tee gpurun_out/r3p_hd_fwd.txt
RFX_G2_TAPGROUPS=0 timeout 300 python tools/hd_bench.py 32 2>&1 | grep HDemucs | tee gpurun_out/r3p_hd_fwd_nogroups.txt
RFX_G2_TAPGROUPS=1 timeout 600 python tools/hd_train_bench.py --batch 16 --steps 4 --warmup 2 > gpurun_out/r3p_hd_train.json 2> gpurun_out/r3p_hd.err; echo "hd train exit=$?"; cut -c1-420 gpurun_out/r3p_hd_train.json

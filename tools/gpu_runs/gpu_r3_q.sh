#!/bin/bash
# (1) gemm2 tap groups, (2) bias column sums inside the fused weight-gradient kernel -- both opt-in: parity first, then timing
mkdir -p gpurun_out
RFX_G2_TAPGROUPS=1 timeout 900 python -m pytest tests/test_gpu_hdemucs.py tests/test_gpu_hdemucs_backward.py tests/test_gpu_tcn.py tests/test_gpu_tcn_backward.py tests/test_gpu_cnn14.py -x -q > gpurun_out/r3q_tests_groups.log 2>&1; echo "tests (tap groups) exit=$?"; tail -4 gpurun_out/r3q_tests_groups.log
RFX_HD_WGRAD_COLSUM=1 timeout 900 python -m pytest tests/test_gpu_hdemucs_backward.py -x -q > gpurun_out/r3q_tests_colsum.log 2>&1; echo "tests (colsum) exit=$?"; tail -4 gpurun_out/r3q_tests_colsum.log
timeout 300 python tools/hd_bench.py 32 2>&1 | grep HDemucs | tee gpurun_out/r3q_hd_fwd.txt
RFX_G2_TAPGROUPS=1 timeout 300 python tools/hd_bench.py 1 16 32 2>&1 | grep HDemucs | tee gpurun_out/r3q_hd_fwd_groups.txt
timeout 600 python tools/hd_train_bench.py --batch 16 --steps 4 --warmup 2 > gpurun_out/r3q_hd_train.json 2> gpurun_out/r3q_hd.err; echo "hd train exit=$?"; cut -c1-330 gpurun_out/r3q_hd_train.json
RFX_G2_TAPGROUPS=1 timeout 600 python tools/hd_train_bench.py --batch 16 --steps 4 --warmup 2 > gpurun_out/r3q_hd_train_groups.json 2> gpurun_out/r3q_hd2.err; echo "hd train (groups) exit=$?"; cut -c1-330 gpurun_out/r3q_hd_train_groups.json
RFX_G2_TAPGROUPS=1 RFX_HD_WGRAD_COLSUM=1 timeout 600 python tools/hd_train_bench.py --batch 16 --steps 4 --warmup 2 > gpurun_out/r3q_hd_train_both.json 2> gpurun_out/r3q_hd3.err; echo "hd train (groups + colsum) exit=$?"; cut -c1-330 gpurun_out/r3q_hd_train_both.json

#!/bin/bash
# training forward: un-normalised layers keep their fused activation epilogue, the same launch stores the pre-activation
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_hdemucs_backward.py tests/test_gpu_hdemucs.py tests/test_gpu_umx.py tests/test_gpu_gemm_lstm.py -x -q > gpurun_out/r3l_tests.log 2>&1; echo "tests exit=$?"; tail -3 gpurun_out/r3l_tests.log
timeout 600 python tools/hd_train_bench.py --batch 16 --steps 4 --warmup 2 > gpurun_out/r3l_hd_train.json 2> gpurun_out/r3l_hd.err; echo "hd train exit=$?"; cut -c1-420 gpurun_out/r3l_hd_train.json
RFX_HD_TRAIN_FUSE_ACT=0 timeout 600 python tools/hd_train_bench.py --batch 16 --steps 4 --warmup 2 > gpurun_out/r3l_hd_train_unfused.json 2> gpurun_out/r3l_hd2.err; echo "hd train (two launches) exit=$?"; cut -c1-420 gpurun_out/r3l_hd_train_unfused.json

#!/bin/bash
# gradient sink: parameter gradients as views of one zeroed flat bucket (no per-parameter memset / accumulate kernels)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_hdemucs_backward.py tests/test_gpu_optim.py tests/test_gpu_zz_train_step.py tests/test_gpu_umx_train.py tests/test_gpu_tcn_backward.py -x -q > gpurun_out/r3j_tests.log 2>&1; echo "tests exit=$?"; tail -3 gpurun_out/r3j_tests.log
timeout 600 python tools/hd_train_bench.py --batch 16 --steps 4 --warmup 2 > gpurun_out/r3j_hd_train.json 2> gpurun_out/r3j_hd.err; echo "hd train exit=$?"; cut -c1-420 gpurun_out/r3j_hd_train.json
timeout 600 python tools/umx_train_bench.py --batch 16 --steps 4 --warmup 2 > gpurun_out/r3j_umx_train.json 2> gpurun_out/r3j_umx.err; echo "umx train exit=$?"; cut -c1-300 gpurun_out/r3j_umx_train.json

#!/bin/bash
# gemm2 with fused GroupNorm statistics: chunked y-fastest tile schedule + running fp64 sums (one flush per segment change)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_hdemucs.py tests/test_gpu_hdemucs_backward.py tests/test_gpu_gemm_lstm.py tests/test_gpu_cnn14.py -x -q > gpurun_out/r3c_tests.log 2>&1; echo "tests exit=$?"; tail -3 gpurun_out/r3c_tests.log
timeout 300 python tools/hd_bench.py 1 16 32 2>&1 | grep HDemucs | tee gpurun_out/r3c_hd_fwd.txt
timeout 600 python tools/hd_train_bench.py --batch 16 --steps 4 --warmup 2 > gpurun_out/r3c_hd_train.json 2> gpurun_out/r3c_hd.err; echo "hd train exit=$?"; cut -c1-420 gpurun_out/r3c_hd_train.json

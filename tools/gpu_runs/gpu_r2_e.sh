#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_tcn_backward.py tests/test_gpu_zz_train_step.py tests/test_gpu_hdemucs_backward.py -m gpu -q --timeout 600 --no-header -p no:cacheprovider > gpurun_out/r2e_tests.log 2>&1
echo "tests exit=$? $(tail -n 1 gpurun_out/r2e_tests.log)"
grep -E "^(FAILED|ERROR)|^E  " gpurun_out/r2e_tests.log | head -20
timeout 600 python tools/tcn_train_bench.py --batch 1 --steps 4 --warmup 2 > gpurun_out/r2e_tcn_train_b1.json 2> gpurun_out/r2e_tcn_train_b1.err
echo "tcn train exit=$?"; cat gpurun_out/r2e_tcn_train_b1.json; tail -n 2 gpurun_out/r2e_tcn_train_b1.err
timeout 600 python tools/hd_train_bench.py --batch 16 --steps 3 --warmup 2 > gpurun_out/r2e_hd_train_b16.json 2> gpurun_out/r2e_hd_train_b16.err
echo "hd train exit=$?"; cat gpurun_out/r2e_hd_train_b16.json; tail -n 2 gpurun_out/r2e_hd_train_b16.err

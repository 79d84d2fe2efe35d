#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_hdemucs_backward.py -m gpu -q --timeout 600 --no-header -p no:cacheprovider -s > gpurun_out/r2l_bwd.log 2>&1
echo "backward exit=$? $(tail -n 1 gpurun_out/r2l_bwd.log)"
grep -E "^(FAILED|ERROR)|^E  |worst weight" gpurun_out/r2l_bwd.log | head -20
timeout 600 python tools/hd_train_bench.py --batch 16 --steps 3 --warmup 2 > gpurun_out/r2l_hd_train_b16.json 2> gpurun_out/r2l_hd_train_b16.err
echo "hd train exit=$?"; cat gpurun_out/r2l_hd_train_b16.json; tail -n 2 gpurun_out/r2l_hd_train_b16.err

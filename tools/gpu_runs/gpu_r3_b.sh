#!/bin/bash
# persistent narrow-layer backward, GroupNorm backward templated on the activation, transposed packs on a preparation stream,
# accumulating input-gradient GEMMs: parity tests of the Hybrid-Demucs backward, then the training step
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_hdemucs_backward.py tests/test_gpu_hdemucs.py tests/test_gpu_gemm_lstm.py -x -q > gpurun_out/r3b_tests.log 2>&1; echo "tests exit=$?"; tail -3 gpurun_out/r3b_tests.log
timeout 600 python tools/hd_train_bench.py --batch 16 --steps 4 --warmup 2 > gpurun_out/r3b_hd_train.json 2> gpurun_out/r3b_hd.err; echo "hd train exit=$?"; cut -c1-600 gpurun_out/r3b_hd_train.json

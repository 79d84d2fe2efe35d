#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_umx_train.py -x -q -s > gpurun_out/r2r_train.log 2>&1; echo "train tests exit=$?"; tail -30 gpurun_out/r2r_train.log
timeout 600 python -m pytest tests/test_gpu_umx.py -x -q > gpurun_out/r2r_umx.log 2>&1; echo "umx tests exit=$?"; tail -5 gpurun_out/r2r_umx.log

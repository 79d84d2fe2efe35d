#!/bin/bash
mkdir -p gpurun_out
N="ncu --set full --clock-control none --import-source on -f"
timeout 300 $N -k regex:stft2048_tma_kernel -s 4 -c 2 -o gpurun_out/r2o_stft_tma python tools/umx_quick_bench.py 32 > gpurun_out/r2o_1.log 2>&1; echo "exit=$?"

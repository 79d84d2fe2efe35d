#!/bin/bash
mkdir -p gpurun_out
N="ncu --set full --clock-control none --import-source on -f"
timeout 600 $N -k regex:gemm2_kernel -s 5 -c 2 -o gpurun_out/r2w_g2_thin python tools/hd_fwd_probe.py 32 > gpurun_out/r2w.log 2>&1; echo "exit=$?"

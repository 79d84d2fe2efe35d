#!/bin/bash
# ncu --set full captures of the last session's backward kernels
mkdir -p gpurun_out
N="ncu --set full --clock-control none --import-source on -f"
timeout 300 $N -k regex:hd_wgrad_tc_fused_kernel -s 0 -c 4 -o gpurun_out/r3o_hd_wgrad_fused python tools/hd_train_bench.py --batch 16 --steps 1 --warmup 0 > gpurun_out/r3o_1.log 2>&1; echo "fused wgrad exit=$?"
timeout 300 $N -k regex:narrow_bwd_kernel -s 0 -c 4 -o gpurun_out/r3o_narrow_bwd python tools/hd_train_bench.py --batch 16 --steps 1 --warmup 0 > gpurun_out/r3o_2.log 2>&1; echo "narrow exit=$?"
timeout 300 $N -k regex:gn_bwd_kernel -s 0 -c 6 -o gpurun_out/r3o_gn_bwd python tools/hd_train_bench.py --batch 16 --steps 1 --warmup 0 > gpurun_out/r3o_3.log 2>&1; echo "gn bwd exit=$?"
ls -la gpurun_out/r3o*.ncu-rep

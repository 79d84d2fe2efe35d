#!/bin/bash
mkdir -p gpurun_out
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2g_launches_hd_train_b16.csv python tools/hd_train_bench.py --batch 16 --steps 1 --warmup 1 > gpurun_out/r2g_ncu.log 2>&1
echo "ncu exit=$?"
python tools/launch_shares.py gpurun_out/r2g_launches_hd_train_b16.csv gpurun_out/r2g_launch_shares_hd_train_b16.txt "second step of tools/hd_train_bench.py --batch 16"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2g_launches_tcn_train_b1.csv python tools/tcn_train_bench.py --batch 1 --steps 1 --warmup 1 > gpurun_out/r2g_ncu_tcn.log 2>&1
python tools/launch_shares.py gpurun_out/r2g_launches_tcn_train_b1.csv gpurun_out/r2g_launch_shares_tcn_train_b1.txt "second step of tools/tcn_train_bench.py --batch 1"

#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r3i_launches_hd_train.csv python tools/hd_train_bench.py --batch 16 --steps 1 --warmup 1 > gpurun_out/r3i_ncu.log 2>&1; echo "ncu hd exit=$?"
python tools/launch_shares.py gpurun_out/r3i_launches_hd_train.csv gpurun_out/r3i_launch_shares_hd_train_b16.txt "second step of tools/hd_train_bench.py --batch 16"
head -45 gpurun_out/r3i_launch_shares_hd_train_b16.txt

#!/bin/bash
# Open-Unmix training step at 16 x 262144 per GPU: stage times + launch shares; fresh launch shares of the Hybrid-Demucs training step
mkdir -p gpurun_out
timeout 600 python tools/umx_train_bench.py --batch 16 --steps 4 --warmup 2 > gpurun_out/r2s_umx_train_b16.json 2> gpurun_out/r2s_umx_train_b16.err; echo "umx train exit=$?"; cat gpurun_out/r2s_umx_train_b16.json; tail -3 gpurun_out/r2s_umx_train_b16.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2s_launches_umx_train.csv python tools/umx_train_bench.py --batch 16 --steps 1 --warmup 1 > gpurun_out/r2s_ncu1.log 2>&1; echo "ncu umx exit=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2s_launches_hd_train.csv python tools/hd_train_bench.py --batch 16 --steps 1 --warmup 1 > gpurun_out/r2s_ncu2.log 2>&1; echo "ncu hd exit=$?"
python tools/launch_shares.py gpurun_out/r2s_launches_umx_train.csv gpurun_out/r2s_launch_shares_umx_train_b16.txt "second step of tools/umx_train_bench.py --batch 16"
python tools/launch_shares.py gpurun_out/r2s_launches_hd_train.csv gpurun_out/r2s_launch_shares_hd_train_b16.txt "second step of tools/hd_train_bench.py --batch 16"
head -22 gpurun_out/r2s_launch_shares_umx_train_b16.txt; head -16 gpurun_out/r2s_launch_shares_hd_train_b16.txt

#!/bin/bash
# ncu --set full captures of the round-2 kernels (one or two launches each)
mkdir -p gpurun_out
N="ncu --set full --clock-control none --import-source on -f"
timeout 200 $N -k regex:tcn_wgrad_tc_kernel -s 9 -c 1 -o gpurun_out/r2m_tcn_wgrad_tc python tools/tcn_train_bench.py --batch 1 --steps 1 --warmup 0 > gpurun_out/r2m_1.log 2>&1; echo "tcn wgrad exit=$?"
timeout 300 $N -k regex:hd_wgrad_tc_kernel -s 0 -c 4 -o gpurun_out/r2m_hd_wgrad_tc python tools/hd_train_bench.py --batch 16 --steps 1 --warmup 0 > gpurun_out/r2m_2.log 2>&1; echo "hd wgrad exit=$?"
timeout 300 $N -k regex:gn_bwd_kernel -s 0 -c 4 -o gpurun_out/r2m_gn_bwd python tools/hd_train_bench.py --batch 16 --steps 1 --warmup 0 > gpurun_out/r2m_3.log 2>&1; echo "gn bwd exit=$?"
timeout 300 $N -k regex:lstm_bwd_persist_kernel -s 0 -c 2 -o gpurun_out/r2m_lstm_bwd_persist python tools/hd_train_bench.py --batch 16 --steps 1 --warmup 0 > gpurun_out/r2m_4.log 2>&1; echo "lstm bwd exit=$?"
timeout 300 $N -k regex:lstm_rec_tc_kernel -s 6 -c 1 -o gpurun_out/r2m_lstm_rec_tc_dual python bench.py --steps 4 --warmup 3 --legs none --no-cpu-baseline --no-gpu-eager > gpurun_out/r2m_5.log 2>&1; echo "lstm rec exit=$?"
ls -la gpurun_out/*.ncu-rep | tail

#!/bin/bash
# TCN training-step measurement on one B200: step timing (B = 1, 4), ncu launch list of one step, ncu --set full of the weight-gradient kernel.
mkdir -p gpurun_out
timeout 100 python tools/tcn_train_bench.py --batch 1 --steps 5 --warmup 2 > gpurun_out/tcn_train_b1.json 2> gpurun_out/tcn_train_b1.err; echo "b1 exit=$?"; cat gpurun_out/tcn_train_b1.json; tail -n 3 gpurun_out/tcn_train_b1.err
timeout 100 python tools/tcn_train_bench.py --batch 4 --steps 2 --warmup 1 > gpurun_out/tcn_train_b4.json 2> gpurun_out/tcn_train_b4.err; echo "b4 exit=$?"; cat gpurun_out/tcn_train_b4.json; tail -n 3 gpurun_out/tcn_train_b4.err
timeout 120 python tools/tcn_train_bench.py --cpu-baseline --steps 2 > gpurun_out/tcn_train_cpu.json 2> gpurun_out/tcn_train_cpu.err; echo "cpu exit=$?"; cat gpurun_out/tcn_train_cpu.json
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_tcn_train.csv python tools/tcn_train_bench.py --batch 1 --steps 1 --warmup 0 > gpurun_out/ncu_tcn_list.log 2>&1; echo "ncu list exit=$?"
timeout 150 ncu --set full --clock-control none --import-source on -k regex:tcn_wgrad_kernel -s 9 -c 1 -f -o gpurun_out/prof_tcn_wgrad python tools/tcn_train_bench.py --batch 1 --steps 1 --warmup 0 > gpurun_out/ncu_tcn_wgrad.log 2>&1; echo "ncu wgrad exit=$?"
ls -la gpurun_out | tail -n 8

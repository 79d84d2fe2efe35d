#!/bin/bash
mkdir -p gpurun_out
python -m remfx_b200.build > /dev/null
timeout 300 python tools/e2e_diag.py 2>&1 | tail -9
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'])"; tail -n 3 gpurun_out/bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches exit=$?"
ncu --set full --clock-control none --import-source on -k regex:lstm_rec -s 12 -c 1 -o gpurun_out/prof_lstm_pipe python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_lstm.log 2>&1; echo "ncu lstm exit=$?"
ncu --set full --clock-control none --import-source on -k regex:stft_kernel -s 8 -c 2 -o gpurun_out/prof_stft_pipe python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_stft.log 2>&1; echo "ncu stft exit=$?"
ls -la gpurun_out | head -30

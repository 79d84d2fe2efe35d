#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_hdemucs_backward.py -m gpu -q --timeout 600 --no-header -p no:cacheprovider -s > gpurun_out/r2d_bwd.log 2>&1
echo "backward exit=$? $(tail -n 1 gpurun_out/r2d_bwd.log)"
grep -E "^(FAILED|ERROR)" gpurun_out/r2d_bwd.log | head
for b in 16; do
  timeout 600 python tools/hd_train_bench.py --batch $b --steps 3 --warmup 2 > gpurun_out/r2d_hd_train_b$b.json 2> gpurun_out/r2d_hd_train_b$b.err
  echo "train b=$b exit=$?"; cat gpurun_out/r2d_hd_train_b$b.json; tail -n 3 gpurun_out/r2d_hd_train_b$b.err
done
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2d_launches_hd_train_b16.csv python tools/hd_train_bench.py --batch 16 --steps 1 --warmup 1 > gpurun_out/r2d_ncu.log 2>&1
echo "ncu exit=$?"
python tools/launch_shares.py gpurun_out/r2d_launches_hd_train_b16.csv gpurun_out/r2d_launch_shares_hd_train_b16.txt "second step of tools/hd_train_bench.py --batch 16"

#!/bin/bash
# gn_bwd grids sized by the instantiation's occupancy (one balanced wave)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_hdemucs_backward.py -x -q > gpurun_out/r3t_tests.log 2>&1; echo "tests exit=$?"; tail -3 gpurun_out/r3t_tests.log
timeout 600 python tools/hd_train_bench.py --batch 16 --steps 4 --warmup 2 > gpurun_out/r3t_hd_train.json 2> gpurun_out/r3t_hd.err; echo "hd train exit=$?"; cut -c1-330 gpurun_out/r3t_hd_train.json

#!/bin/bash
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_green.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_green.log 2>&1; echo "ncu launches (green) exit=$?"; tail -5 gpurun_out/ncu_launch_green.log
RFX_UMX_PIPE_GREEN=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches (caps) exit=$?"; tail -3 gpurun_out/ncu_launch.log | cut -c1-300

#!/bin/bash
mkdir -p gpurun_out
RFX_G2_TAPGROUPS=2 timeout 900 python -m pytest tests/test_gpu_hdemucs.py tests/test_gpu_hdemucs_backward.py tests/test_gpu_tcn.py tests/test_gpu_tcn_backward.py tests/test_gpu_cnn14.py -x -q > gpurun_out/r3s_tests_groups.log 2>&1; echo "tests (tap groups, mode 2) exit=$?"; tail -3 gpurun_out/r3s_tests_groups.log
for m in 0 2; do
RFX_G2_TAPGROUPS=$m RFX_G2_TRACE=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:gemm2_kernel --csv --log-file gpurun_out/r3s_g2_launches_m$m.csv python tools/hd_fwd_probe.py 32 > gpurun_out/r3s_out_m$m.log 2> gpurun_out/r3s_trace_m$m.log; echo "mode $m exit=$?"
done

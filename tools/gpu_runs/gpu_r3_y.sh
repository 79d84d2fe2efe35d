#!/bin/bash
# compute-sanitizer racecheck on gemm2 (staged fp32 rows: st.shared / __syncwarp / ld.shared per warp) and on the fused weight gradient
mkdir -p gpurun_out
CS="/usr/local/cuda/bin/compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 10"
timeout 900 $CS --kernel-regex kns=gemm2 python -m pytest "tests/test_gpu_hdemucs.py::test_layerwise_and_output" -x -q -k "over0" > gpurun_out/r3y_race_gemm2.log 2>&1; echo "racecheck gemm2 exit=$?"; grep "RACECHECK SUMMARY\|passed\|failed\|hazard" gpurun_out/r3y_race_gemm2.log | tail -5
timeout 900 $CS --kernel-regex kns=wgrad_tc_fused python -m pytest "tests/test_gpu_hdemucs_backward.py" -x -q -k "conv_only" > gpurun_out/r3y_race_wgrad.log 2>&1; echo "racecheck wgrad exit=$?"; grep "RACECHECK SUMMARY\|passed\|failed\|hazard" gpurun_out/r3y_race_wgrad.log | tail -5

#!/bin/bash
mkdir -p gpurun_out
for m in 3 2 1; do
RFX_G2_TAPGROUPS=$m timeout 300 python -m pytest "tests/test_gpu_hdemucs.py::test_layerwise_and_output" -x -q > gpurun_out/r3r_tests_m$m.log 2>&1; echo "mode $m exit=$?"; grep -n "rel-RMS" gpurun_out/r3r_tests_m$m.log | head -8; tail -2 gpurun_out/r3r_tests_m$m.log
done

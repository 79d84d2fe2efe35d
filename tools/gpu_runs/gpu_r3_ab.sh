#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_hdemucs_backward.py -x -q -k "off_the_hop_grid" > gpurun_out/r3ab_tests.log 2>&1; echo "tests exit=$?"; tail -15 gpurun_out/r3ab_tests.log

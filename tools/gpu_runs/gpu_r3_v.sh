#!/bin/bash
# Hybrid-Demucs inference on lengths off the 1024-sample grid
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_hdemucs.py -x -q > gpurun_out/r3v_tests.log 2>&1; echo "tests exit=$?"; grep -n "rel-RMS\|Error\|error" gpurun_out/r3v_tests.log | tail -12; tail -3 gpurun_out/r3v_tests.log

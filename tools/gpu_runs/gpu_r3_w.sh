#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_chain.py -x -q > gpurun_out/r3w_tests.log 2>&1; echo "tests exit=$?"; tail -15 gpurun_out/r3w_tests.log

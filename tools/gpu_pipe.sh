#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_umx.py -m gpu -q --timeout 120 --no-header -p no:cacheprovider > gpurun_out/t.log 2>&1; echo "umx tests exit=$? $(tail -n 1 gpurun_out/t.log)"
timeout 300 python tools/e2e_diag.py 2>&1 | tail -8

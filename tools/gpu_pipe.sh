#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_umx.py -m gpu -q --timeout 200 --no-header -p no:cacheprovider > gpurun_out/t.log 2>&1; echo "umx tests exit=$? $(tail -n 1 gpurun_out/t.log)"; grep -E "^FAILED|^ERROR|rror" gpurun_out/t.log | head
RFX_UMX_PIPE_GREEN=0 timeout 600 python -m pytest tests/test_gpu_umx.py -m gpu -q --timeout 200 --no-header -p no:cacheprovider -k pipeline > gpurun_out/t2.log 2>&1; echo "umx pipeline tests (grid caps) exit=$? $(tail -n 1 gpurun_out/t2.log)"

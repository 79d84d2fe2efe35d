#!/bin/bash
mkdir -p gpurun_out
for f in tests/test_gpu_stft.py tests/test_gpu_umx.py tests/test_gpu_cnn14.py tests/test_gpu_hdemucs.py; do
timeout 600 python -m pytest $f -m gpu -q --timeout 200 --no-header -p no:cacheprovider > gpurun_out/t.log 2>&1; echo "$f exit=$? $(tail -n 1 gpurun_out/t.log)"; grep -E "^FAILED|^ERROR|rror:" gpurun_out/t.log | head -8
done
timeout 120 python tools/umx_quick_bench.py 32 2>&1 | tail -2
for i in 1 2; do timeout 120 python tools/pipe_bench.py 32 40 2>&1 | tail -1 | cut -c1-120; done

#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_umx.py tests/test_gpu_ingest.py -m gpu -q --timeout 200 --no-header -p no:cacheprovider > gpurun_out/t.log 2>&1; echo "umx tests exit=$? $(tail -n 1 gpurun_out/t.log)"; grep -E "^FAILED|^ERROR|rror|assert" gpurun_out/t.log | head -8
RFX_UMX_PIPE_LANES=3 timeout 600 python -m pytest tests/test_gpu_umx.py -m gpu -q --timeout 200 --no-header -p no:cacheprovider -k pipeline > gpurun_out/t2.log 2>&1; echo "staggered-mode pipeline tests exit=$? $(tail -n 1 gpurun_out/t2.log)"
for cfg in "6 32 3" "5 32 3" "8 32 3" "6 16 2" "6 16 3" "3 16 2"; do set -- $cfg
RFX_UMX_PIPE_LANES=$1 RFX_UMX_PIPE_SLOTS=$2 RFX_UMX_PIPE_REC_STREAMS=$3 timeout 120 python tools/pipe_bench.py 32 60 2>&1 | tail -1 | cut -c1-230 | sed "s/^/lanes=$1 /"
done

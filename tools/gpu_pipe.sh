#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_umx.py tests/test_gpu_stft.py tests/test_gpu_gemm_lstm.py -m gpu -q --timeout 300 --no-header -p no:cacheprovider > gpurun_out/test_pipe.log 2>&1; echo "tests exit=$? $(tail -n 1 gpurun_out/test_pipe.log)"
grep -E "FAILED|Error|error" gpurun_out/test_pipe.log | head -20
timeout 300 python tools/pipe_bench.py 32 40 2>&1 | tail -3
RFX_UMX_PIPE_MAX_SMS=0 timeout 300 python tools/pipe_bench.py 32 40 2>&1 | tail -1
RFX_UMX_PIPE_MAX_SMS=0 RFX_UMX_PIPE_SLOTS=0 timeout 300 python tools/pipe_bench.py 32 40 2>&1 | tail -1
RFX_UMX_PIPE_MAX_SMS=100 timeout 300 python tools/pipe_bench.py 32 40 2>&1 | tail -1
RFX_UMX_PIPE_MAX_SMS=70 timeout 300 python tools/pipe_bench.py 32 40 2>&1 | tail -1
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit=$?"; cat gpurun_out/bench.json; tail -n 3 gpurun_out/bench.err

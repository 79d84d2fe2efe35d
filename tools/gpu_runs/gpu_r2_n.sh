#!/bin/bash
# staged (bulk-copy) STFT / iSTFT for n_fft = 2048: parity tests, A/B against the legacy kernels, pipeline with 6 / 8 lanes
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_stft.py tests/test_gpu_umx.py tests/test_gpu_full_size.py tests/test_gpu_cnn14.py -x -q > gpurun_out/r2n_tests.log 2>&1; echo "tests exit=$?"; tail -5 gpurun_out/r2n_tests.log
echo "== staged"; timeout 300 python tools/umx_quick_bench.py 32 2>&1 | tail -3
echo "== legacy"; RFX_STFT_LEGACY=1 timeout 300 python tools/umx_quick_bench.py 32 2>&1 | tail -3
for h in 13 16 21; do echo "== staged hops=$h"; RFX_ISTFT_HOPS=$h timeout 300 python tools/umx_quick_bench.py 32 2>&1 | tail -1; done
echo "== pipe staged 6 lanes"; timeout 300 python tools/pipe_bench.py 32 200 2>&1 | tail -1
echo "== pipe staged 8 lanes"; RFX_UMX_PIPE_LANES=8 timeout 300 python tools/pipe_bench.py 32 200 2>&1 | tail -1
echo "== pipe legacy 6 lanes"; RFX_STFT_LEGACY=1 timeout 300 python tools/pipe_bench.py 32 200 2>&1 | tail -1
echo "== pipe legacy 8 lanes"; RFX_STFT_LEGACY=1 RFX_UMX_PIPE_LANES=8 timeout 300 python tools/pipe_bench.py 32 200 2>&1 | tail -1

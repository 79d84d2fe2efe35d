#!/bin/bash
# compute-sanitizer memcheck over the new code paths at small sizes (any-length forward, staged gemm2 stores, fused weight gradient,
# persistent narrow-layer backward, gradient sink).  The cluster recurrence kernels are excluded: memcheck reports their DSMEM bulk
# copies (cp.async.bulk.shared::cluster into a PEER CTA's shared memory) as invalid shared writes.
mkdir -p gpurun_out
CS="/usr/local/cuda/bin/compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 30 --kernel-regex-exclude kns=lstm"
timeout 900 $CS python -m pytest "tests/test_gpu_hdemucs.py::test_any_length_matches_torchaudio" -x -q -k "4103 or 20000" > gpurun_out/r3x_memcheck_fwd.log 2>&1; echo "memcheck fwd exit=$?"; grep -c "Invalid\|out of bounds\|misaligned" gpurun_out/r3x_memcheck_fwd.log; grep "ERROR SUMMARY\|passed\|failed" gpurun_out/r3x_memcheck_fwd.log | tail -3
timeout 1200 $CS python -m pytest "tests/test_gpu_hdemucs_backward.py" -x -q -k "linear_objective" > gpurun_out/r3x_memcheck_bwd.log 2>&1; echo "memcheck bwd exit=$?"; grep -c "Invalid\|out of bounds\|misaligned" gpurun_out/r3x_memcheck_bwd.log; grep "ERROR SUMMARY\|passed\|failed" gpurun_out/r3x_memcheck_bwd.log | tail -3
timeout 1200 $CS python -m pytest "tests/test_gpu_hdemucs_backward.py" -x -q -k "fit_step" > gpurun_out/r3x_memcheck_fit.log 2>&1; echo "memcheck fit exit=$?"; grep -c "Invalid\|out of bounds\|misaligned" gpurun_out/r3x_memcheck_fit.log; grep "ERROR SUMMARY\|passed\|failed" gpurun_out/r3x_memcheck_fit.log | tail -3

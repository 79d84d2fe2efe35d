#!/bin/bash
for h in 8 16 24 32; do
RFX_ISTFT_HPC=$h timeout 120 python tools/umx_quick_bench.py 32 2>&1 | tail -1 | sed "s/^/hpc=$h /"
RFX_ISTFT_HPC=$h timeout 120 python tools/pipe_bench.py 32 40 2>&1 | tail -1 | cut -c1-40 | sed "s/^/hpc=$h /"
done

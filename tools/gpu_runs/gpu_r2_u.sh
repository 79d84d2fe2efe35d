#!/bin/bash
mkdir -p gpurun_out
RFX_G2_TRACE=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:gemm2_kernel --csv --log-file gpurun_out/r2u_g2_launches.csv python tools/hd_fwd_probe.py 32 > gpurun_out/r2u_out.log 2> gpurun_out/r2u_trace.log; echo "exit=$?"
grep -c g2trace gpurun_out/r2u_trace.log

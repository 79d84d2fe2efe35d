#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/lstm_tc_probe.py 2>&1 | tail -3
timeout 120 python tools/lstm_bench.py 32 2>&1 | tail -2
RFX_LSTM_IMPL=2 RFX_UMX_PIPE_SLOTS=16 timeout 120 python tools/pipe_bench.py 32 40 2>&1 | tail -2
RFX_LSTM_IMPL=2 RFX_UMX_PIPE_SLOTS=16 RFX_UMX_PIPE_REC_STREAMS=2 timeout 120 python tools/pipe_bench.py 32 40 2>&1 | tail -1

#!/bin/bash
for d in 96 288 96 288; do echo "== RFX_LSTM_TC_DEBUG=$d"; RFX_LSTM_TC_DEBUG=$d timeout 100 python tools/lstm_bench.py 32 2>&1 | grep "slots=32"; done

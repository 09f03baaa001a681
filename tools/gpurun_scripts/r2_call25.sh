#!/bin/bash
# Round 2, call 25: windowed attention, loads only (TMA boxes of 196 x 32 B rows; no MMA, no softmax) vs the full kernel
mkdir -p gpurun_out
timeout 120 python tools/attention_probe.py 8 fp16 2>&1 | tee gpurun_out/r2c25_probe_full.log
B200SAM_WIN_LOADS_ONLY=1 timeout 120 python tools/attention_probe.py 8 fp16 2>&1 | tee gpurun_out/r2c25_probe_loads_only.log

#!/bin/bash
# Round 2, call 12: per-CTA timeline of the whole-row windowed attention kernel
mkdir -p gpurun_out
timeout 120 python tools/window_attention_timeline.py 8 2>&1 | tee gpurun_out/r2c12_timeline.log

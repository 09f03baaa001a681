#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/profile_attention.py 2 > gpurun_out/pa_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:window_attn_tc -s 1 -c 1 -f -o gpurun_out/win_r1 python tools/profile_attention.py 2 > gpurun_out/ncu_win.log 2>&1
echo "ncu exit=$?"
timeout 300 python tools/attn_bench.py 8 2>&1 | tail -5

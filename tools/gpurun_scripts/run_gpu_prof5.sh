#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/profile_attention2.py 8 4 > gpurun_out/pa2_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:window_attn_tc2 -s 2 -c 1 -f -o gpurun_out/win2_r1 python tools/profile_attention2.py 8 4 > gpurun_out/ncu_win2.log 2>&1
echo "ncu exit=$?"

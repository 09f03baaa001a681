#!/bin/bash
# Round 2, call 9: proj epilogue variants (probe), attention ncu captures for profiles/
mkdir -p gpurun_out
timeout 180 python tools/gemm_pair_probe.py 20 > gpurun_out/r2c9_probe.log 2>&1; echo "probe exit=$?"; cat gpurun_out/r2c9_probe.log
timeout 120 python tools/profile_attention.py 8 fp16 > gpurun_out/r2c9_attn_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_tc -s 2 -c 2 -f -o gpurun_out/r2c9_attn python tools/profile_attention.py 8 fp16 > gpurun_out/r2c9_ncu_attn.log 2>&1
echo "ncu attn exit=$?"

#!/bin/bash
# Round 2, call 34: ncu --set full of the decoder's fused k | v | q projection GEMM (tall tiles, 3 planes, M = 120 x 4096)
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_bf16_tn_kernel --launch-skip 2 --launch-count 1 -f -o gpurun_out/r2c34_gemm_planes python tools/profile_decode_stage.py 8 all > gpurun_out/r2c34_ncu.log 2>&1
echo "ncu exit=$?"; tail -2 gpurun_out/r2c34_ncu.log

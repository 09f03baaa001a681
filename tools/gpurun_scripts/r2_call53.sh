#!/bin/bash
# Round 2, call 53: ncu --set full of the decoder's fp32 attention kernels (image->token, token->image) inside one refinement
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:attn_few --launch-skip 4 --launch-count 4 -f -o gpurun_out/r2c53_dec_attn python tools/profile_decode_stage.py 8 all > gpurun_out/r2c53_ncu.log 2>&1
echo "ncu exit=$?"; tail -2 gpurun_out/r2c53_ncu.log

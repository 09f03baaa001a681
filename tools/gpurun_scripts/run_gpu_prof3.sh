#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/profile_decode_stage.py 8 stages > gpurun_out/pds2_plain.log 2>&1 && \
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:'upscale_mask_fast|prompt_accum' -f -o gpurun_out/stages_r2 python tools/profile_decode_stage.py 8 stages > gpurun_out/ncu_pds2.log 2>&1
echo "ncu full exit=$?"

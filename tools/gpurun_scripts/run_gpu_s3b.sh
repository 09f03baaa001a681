#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/profile_decode_stage.py 8 stages > gpurun_out/pds_plain.log 2>&1; echo "plain exit=$?"
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:'upscale_mask_fast|prompt_accum' -f -o gpurun_out/stages_s3 python tools/profile_decode_stage.py 8 stages > gpurun_out/ncu_pds.log 2>&1; echo "ncu exit=$?"

#!/bin/bash
# Round 2, call 31: launch list of the decode stage (8 images, 120 prompts, 2 passes) at the end of round 2
mkdir -p gpurun_out
timeout 300 python tools/profile_decode_stage.py 8 all > gpurun_out/r2c31_decode_plain.log 2>&1 && \
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2c31_launches_decode_stage.csv python tools/profile_decode_stage.py 8 all > gpurun_out/r2c31_ncu_decode.log 2>&1
echo "ncu decode exit=$?"; tail -3 gpurun_out/r2c31_decode_plain.log

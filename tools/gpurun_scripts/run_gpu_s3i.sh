#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/profile_unet.py 8 > gpurun_out/pu_plain.log 2>&1 && \
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_unet.csv python tools/profile_unet.py 8 > gpurun_out/ncu_unet.log 2>&1
echo "ncu unet exit=$?"; tail -2 gpurun_out/pu_plain.log
python tools/summarize_launches.py gpurun_out/launches_unet.csv | head -30

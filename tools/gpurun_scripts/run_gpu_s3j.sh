#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_unet_gpu.py -m gpu -q -x --tb=short 2>&1 | tail -3
timeout 300 python tools/profile_unet.py 8 > gpurun_out/pu_plain.log 2>&1 && \
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_unet.csv python tools/profile_unet.py 8 > gpurun_out/ncu_unet.log 2>&1
echo "ncu unet exit=$?"
python tools/summarize_launches.py gpurun_out/launches_unet.csv | head -14

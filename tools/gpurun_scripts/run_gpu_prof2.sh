#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/profile_decode_stage.py 8 all > gpurun_out/pds_plain.log 2>&1 && \
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_decode_stage.csv python tools/profile_decode_stage.py 8 all > gpurun_out/ncu_pds.log 2>&1
echo "ncu launches exit=$?"
python tools/summarize_launches.py gpurun_out/launches_decode_stage.csv
timeout 300 python tools/profile_decode_stage.py 8 stages > gpurun_out/pds2_plain.log 2>&1 && \
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:'upscale_mask_fast|nearest_from_mask|prompt_accum' -f -o gpurun_out/stages_r1 python tools/profile_decode_stage.py 8 stages > gpurun_out/ncu_pds2.log 2>&1
echo "ncu full exit=$?"

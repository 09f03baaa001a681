#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit=$?"; tail -12 gpurun_out/pytest_gpu.log
timeout 300 python tools/stage_bench.py > gpurun_out/stage.json 2> gpurun_out/stage.err; echo "stage exit=$?"; grep -E '"ms"|gbs|upscale|prompt' gpurun_out/stage.json
timeout 300 python tools/profile_decode_stage.py 8 refine > gpurun_out/pds_plain.log 2>&1 && \
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_decode_stage.csv python tools/profile_decode_stage.py 8 refine > gpurun_out/ncu_pds.log 2>&1
echo "ncu launches exit=$?"
python tools/summarize_launches.py gpurun_out/launches_decode_stage.csv | head -24

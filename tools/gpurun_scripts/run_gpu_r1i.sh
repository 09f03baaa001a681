#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -q --tb=short -x -k "linear" > gpurun_out/pytest_lin.log 2>&1; echo "pytest exit=$?"; tail -3 gpurun_out/pytest_lin.log | cut -c1-300
timeout 900 python -m pytest tests/test_model_gpu.py -m gpu -q --tb=short -x -k "refine or predict or pipeline" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit=$?"; tail -4 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 300 python tools/profile_decode_stage.py 8 refine > gpurun_out/pds_plain.log 2>&1 && \
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_decode_stage.csv python tools/profile_decode_stage.py 8 refine > gpurun_out/ncu_pds.log 2>&1
echo "ncu launches exit=$?"
python tools/summarize_launches.py gpurun_out/launches_decode_stage.csv | head -12; python tools/summarize_launches.py gpurun_out/launches_decode_stage.csv | tail -1

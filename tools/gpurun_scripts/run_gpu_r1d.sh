#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py -m gpu -q --tb=short -x -k "upscale or prompt" > gpurun_out/pytest_ops.log 2>&1; echo "pytest exit=$?"; tail -5 gpurun_out/pytest_ops.log
timeout 300 python tools/stage_bench.py > gpurun_out/stage.json 2> gpurun_out/stage.err; echo "stage exit=$?"; grep -E '"ms"|gbs|frac|upscale|prompt' gpurun_out/stage.json; tail -3 gpurun_out/stage.err
timeout 300 python tools/profile_decode_stage.py 8 stages > gpurun_out/pds2_plain.log 2>&1 && \
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:'upscale_mask_fast|prompt_accum' -f -o gpurun_out/stages_r3 python tools/profile_decode_stage.py 8 stages > gpurun_out/ncu_pds2.log 2>&1
echo "ncu full exit=$?"

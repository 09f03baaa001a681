#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py -m gpu -q --tb=short -x -k "upscale or prompt" > gpurun_out/pytest_ops.log 2>&1; echo "pytest exit=$?"; tail -15 gpurun_out/pytest_ops.log
timeout 300 python tools/stage_bench.py > gpurun_out/stage.json 2> gpurun_out/stage.err; echo "stage exit=$?"; cat gpurun_out/stage.json | head -40; tail -3 gpurun_out/stage.err

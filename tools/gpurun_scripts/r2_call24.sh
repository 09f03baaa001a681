#!/bin/bash
# Round 2, call 24: windowed attention requests its tiles before the CTA-wide set-up; idle warps skipped
mkdir -p gpurun_out
timeout 120 python tools/attention_probe.py 8 fp16 2>&1 | tee gpurun_out/r2c24_probe.log
timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -x -q --tb=short -k "attention" > gpurun_out/r2c24_pytest_att.log 2>&1; echo "pytest attention exit=$?"; tail -3 gpurun_out/r2c24_pytest_att.log

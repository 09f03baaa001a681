#!/bin/bash
# Round 2, call 23: windowed attention with the PV wait ahead of the exponentials + one pair in four on the FMA pipe + idle warps skipped
mkdir -p gpurun_out
timeout 120 python tools/attention_probe.py 8 fp16 2>&1 | tee gpurun_out/r2c23_probe.log
timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -x -q --tb=short -k "attention" > gpurun_out/r2c23_pytest_att.log 2>&1; echo "pytest attention exit=$?"; tail -3 gpurun_out/r2c23_pytest_att.log

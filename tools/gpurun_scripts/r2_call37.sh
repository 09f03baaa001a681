#!/bin/bash
# Round 2, call 37: new persistent-grid shape test + the attention tests
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -x -q --tb=short -k "attention" > gpurun_out/r2c37_pytest_att.log 2>&1; echo "pytest attention exit=$?"; tail -3 gpurun_out/r2c37_pytest_att.log

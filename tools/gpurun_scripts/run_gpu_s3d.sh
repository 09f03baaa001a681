#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_resize_gpu.py tests/test_model_gpu.py -m gpu -q --tb=short -x > gpurun_out/pytest_resize.log 2>&1; echo "pytest exit=$?"; tail -15 gpurun_out/pytest_resize.log | cut -c1-300

#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -q --tb=short -x -k "attention" > gpurun_out/pytest_attn.log 2>&1; echo "pytest exit=$?"; tail -7 gpurun_out/pytest_attn.log | cut -c1-300
timeout 200 python tools/attn_bench.py 8 2>&1 | tail -6

#!/bin/bash
# Round 2, call 42: full GPU test suite after the decoder / GEMM changes (pair kernel: a_wrap + planes; direct epilogue removed)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --tb=short -s > gpurun_out/r2c42_pytest.log 2>&1; echo "pytest exit=$?"; grep -E "passed|failed|error" gpurun_out/r2c42_pytest.log | tail -3
grep -h "mismatched\|min dice\|^e2e" gpurun_out/r2c42_pytest.log | head -12
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2c42_smoke.log 2>&1; echo "smoke exit=$?"; tail -1 gpurun_out/r2c42_smoke.log

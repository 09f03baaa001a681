#!/bin/bash
# Round 2, call 32: decode-stage launch list; compute-sanitizer memcheck over the attention and GEMM op tests
mkdir -p gpurun_out
timeout 300 python tools/profile_decode_stage.py 8 all > gpurun_out/r2c31_decode_plain.log 2>&1 && \
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2c31_launches_decode_stage.csv python tools/profile_decode_stage.py 8 all > gpurun_out/r2c31_ncu_decode.log 2>&1
echo "ncu decode exit=$?"; tail -3 gpurun_out/r2c31_decode_plain.log
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 1 --log-file gpurun_out/r2c32_memcheck_attention.log python -m pytest tests/test_ops_gpu.py -m gpu -x -q -k "attention" > gpurun_out/r2c32_memcheck_pytest.log 2>&1; echo "memcheck attention exit=$?"; tail -2 gpurun_out/r2c32_memcheck_pytest.log; grep -c "Invalid\|Error" gpurun_out/r2c32_memcheck_attention.log; tail -3 gpurun_out/r2c32_memcheck_attention.log

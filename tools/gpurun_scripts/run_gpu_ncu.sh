#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/profile_gemm.py > gpurun_out/pg_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 4 -c 4 -f -o gpurun_out/gemm_r1 python tools/profile_gemm.py > gpurun_out/ncu_gemm.log 2>&1
echo "ncu exit=$?"; tail -3 gpurun_out/ncu_gemm.log; ls -la gpurun_out/*.ncu-rep

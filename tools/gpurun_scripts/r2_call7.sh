#!/bin/bash
# Round 2, call 7: ncu --set full with source counters on the pair GEMM (proj, qkv, lin1, lin2 of the probe) + launch list of one encoder forward
mkdir -p gpurun_out
timeout 120 python tools/gemm_pair_probe.py 1 > gpurun_out/r2c7_probe_plain.log 2>&1; echo "probe exit=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_pair -c 17 -f -o gpurun_out/r2c7_gemm_pair python tools/gemm_pair_probe.py 1 > gpurun_out/r2c7_ncu_pair.log 2>&1; echo "ncu pair exit=$?"
timeout 300 python tools/profile_encoder.py --batch 8 --iters 2 > gpurun_out/r2c7_prof_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:b200sam -s 167 -c 167 --csv --log-file gpurun_out/r2c7_launches_enc_b8.csv python tools/profile_encoder.py --batch 8 --iters 2 > gpurun_out/r2c7_ncu_enc.log 2>&1
echo "ncu launches exit=$?"; tail -2 gpurun_out/r2c7_ncu_enc.log

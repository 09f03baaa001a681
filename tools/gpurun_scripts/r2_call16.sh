#!/bin/bash
# Round 2, call 16: global attention with P handed to the PV MMA through TMEM (A operand in TMEM) vs through shared memory
mkdir -p gpurun_out
B200SAM_GLOBATTN=smem timeout 120 python tools/attention_probe.py 8 fp16 save gpurun_out/r2c16_att.pt 2>&1 | tee gpurun_out/r2c16_probe_smem.log
timeout 120 python tools/attention_probe.py 8 fp16 check gpurun_out/r2c16_att.pt 2>&1 | tee gpurun_out/r2c16_probe_tmem.log
rm -f gpurun_out/r2c16_att.pt
timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -x -q --tb=short -k "attention" > gpurun_out/r2c16_pytest_att.log 2>&1; echo "pytest attention exit=$?"; tail -3 gpurun_out/r2c16_pytest_att.log

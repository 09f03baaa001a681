#!/bin/bash
# Round 2, call 27: persistent windowed attention kernel (B200SAM_WINATTN=persist) vs the per-window CTA kernel
mkdir -p gpurun_out
timeout 120 python tools/attention_probe.py 8 fp16 save gpurun_out/r2c27_att.pt 2>&1 | tee gpurun_out/r2c27_probe_cta.log
B200SAM_WINATTN=persist timeout 90 python tools/attention_probe.py 8 fp16 check gpurun_out/r2c27_att.pt 2>&1 | tee gpurun_out/r2c27_probe_persist.log; echo "persist probe exit=${PIPESTATUS[0]}"
rm -f gpurun_out/r2c27_att.pt
B200SAM_WINATTN=persist timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -x -q --tb=short -k "attention" > gpurun_out/r2c27_pytest_att.log 2>&1; echo "pytest attention (persist) exit=$?"; tail -3 gpurun_out/r2c27_pytest_att.log

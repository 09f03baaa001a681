#!/bin/bash
# Round 2, call 22: global attention with the PV wait moved ahead of the exponentials (MUFU interleaved with pack / sums), both row-sum variants
# softmax threads; windowed kernel with inactive warps skipped
mkdir -p gpurun_out
B200SAM_GLOBATTN=sums timeout 120 python tools/attention_probe.py 8 fp16 save gpurun_out/r2c22_att.pt 2>&1 | tee gpurun_out/r2c22_probe_sums.log
timeout 120 python tools/attention_probe.py 8 fp16 check gpurun_out/r2c22_att.pt 2>&1 | tee gpurun_out/r2c22_probe_diet.log
rm -f gpurun_out/r2c22_att.pt
timeout 120 python tools/attention_probe.py 8 bf16 2>&1 | tee gpurun_out/r2c22_probe_diet_bf16.log
timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -x -q --tb=short -k "attention" > gpurun_out/r2c22_pytest_att.log 2>&1; echo "pytest attention exit=$?"; tail -3 gpurun_out/r2c22_pytest_att.log
timeout 900 python -m pytest tests/test_model_gpu.py -m gpu -x -q --tb=short > gpurun_out/r2c22_pytest_model.log 2>&1; echo "pytest model exit=$?"; tail -3 gpurun_out/r2c22_pytest_model.log
grep -h "rel_l2\|Dice\|dice" gpurun_out/r2c22_pytest_model.log | head

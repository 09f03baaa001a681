#!/bin/bash
# One gpurun call: op-level parity first (each group under its own timeout), then the model-level tests.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
for t in prompt_extraction upscale layernorm linear_f32 gemm encoder_attention; do
  echo "=== $t" >> gpurun_out/ops.log
  timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -q -k "$t" -x --tb=short >> gpurun_out/ops.log 2>&1
  echo "exit=$?" >> gpurun_out/ops.log
done
tail -5 gpurun_out/ops.log
timeout 1200 python -m pytest tests/test_model_gpu.py -m gpu -q -s --tb=short > gpurun_out/model.log 2>&1
echo "model exit=$?" >> gpurun_out/model.log
tail -30 gpurun_out/model.log

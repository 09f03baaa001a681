#!/bin/bash
# Round 2, call 14: global attention with 8 x 8 key blocks (8 + 8 bias terms per row and tile instead of 64 + 1)
mkdir -p gpurun_out
timeout 120 python tools/attention_probe.py 8 fp16 2>&1 | tee gpurun_out/r2c14_probe.log
timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -x -q --tb=short -k "attention" > gpurun_out/r2c14_pytest_att.log 2>&1; echo "pytest attention exit=$?"; tail -3 gpurun_out/r2c14_pytest_att.log
timeout 600 python -m pytest tests/test_model_gpu.py -m gpu -x -q --tb=short > gpurun_out/r2c14_pytest_model.log 2>&1; echo "pytest model exit=$?"; tail -3 gpurun_out/r2c14_pytest_model.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-refine --no-cpu-baseline > gpurun_out/r2c14_bench.json 2> gpurun_out/r2c14_bench.err
echo "bench exit=$? $(python -c "
import json;d=json.load(open('gpurun_out/r2c14_bench.json'));r=d['roofline']
print(round(d['value'],2), round(d['ms_per_step'],3), d['clocks'].get('sm_mhz'), 'gemmTF', round(r['achieved'],1), {k:v['ms_mean'] for k,v in r['attention'].items()})" 2>&1)"

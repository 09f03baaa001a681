#!/bin/bash
# Round 2, call 6 (1 GPU): f24 residual stream; A/B against the fp32 stream; ncu of the fp32-output GEMMs
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_model_gpu.py -m gpu -x -q --tb=short -s > gpurun_out/r2c6_pytest.log 2>&1; echo "pytest exit=$?"; grep -E "passed|failed|error" gpurun_out/r2c6_pytest.log | tail -3; grep -E "^e2e|^encoder vit_b|^encoder vit_h" gpurun_out/r2c6_pytest.log
for cfg in "B200SAM_RESIDUAL=f24" "B200SAM_RESIDUAL=fp32" "B200SAM_RESIDUAL=f24 B200SAM_PDL=0"; do
  tag=$(echo "$cfg" | tr ' =' '__')
  env $cfg timeout 300 python bench.py --steps 10 --warmup 3 --no-refine --no-cpu-baseline > gpurun_out/r2c6_bench_$tag.json 2> gpurun_out/r2c6_bench_$tag.err
  echo "$cfg exit=$? $(python -c "
import json;d=json.load(open('gpurun_out/r2c6_bench_$tag.json'));r=d['roofline']
print(round(d['value'],2), round(d['ms_per_step'],3), d['clocks'].get('sm_mhz'), 'gemmTF', round(r['achieved'],1), {k:v['ms_mean'] for k,v in r['per_shape'].items() if k in ('qkv','proj','lin1','lin2')}, {k:v['ms_mean'] for k,v in r['attention'].items()})" 2>&1)"
done
tail -3 gpurun_out/r2c6_bench_B200SAM_RESIDUAL_f24.err

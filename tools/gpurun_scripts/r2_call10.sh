#!/bin/bash
# Round 2, call 10: direct (un-staged) fp32 epilogue of the pair GEMM
mkdir -p gpurun_out
timeout 240 python tools/gemm_pair_probe.py 20 > gpurun_out/r2c10_probe.log 2>&1; echo "probe exit=$?"; cat gpurun_out/r2c10_probe.log
timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_model_gpu.py -m gpu -x -q --tb=short > gpurun_out/r2c10_pytest.log 2>&1; echo "pytest exit=$?"; tail -2 gpurun_out/r2c10_pytest.log
for cfg in "B200SAM_GEMM_DIRECT=1" "B200SAM_GEMM_DIRECT=0" "B200SAM_GEMM_DIRECT=1 B200SAM_ENCODER_OPERANDS=bf16"; do
  tag=$(echo "$cfg" | tr ' =' '__')
  env $cfg timeout 300 python bench.py --steps 10 --warmup 3 --no-refine --no-cpu-baseline > gpurun_out/r2c10_bench_$tag.json 2> gpurun_out/r2c10_bench_$tag.err
  echo "$cfg exit=$? $(python -c "
import json;d=json.load(open('gpurun_out/r2c10_bench_$tag.json'));r=d['roofline']
print(round(d['value'],2), round(d['ms_per_step'],3), d['clocks'].get('sm_mhz'), 'gemmTF', round(r['achieved'],1), {k:v['ms_mean'] for k,v in r['per_shape'].items() if k in ('qkv','proj','lin1','lin2')}, {k:v['ms_mean'] for k,v in r['attention'].items()})" 2>&1)"
done

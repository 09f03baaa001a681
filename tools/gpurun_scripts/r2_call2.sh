#!/bin/bash
# Round 2, call 2: CTA-pair GEMM probe + parity + bench A/B + ncu of the pair kernel
mkdir -p gpurun_out
timeout 180 python tools/gemm_pair_probe.py 20 > gpurun_out/r2c2_probe.log 2>&1; echo "probe exit=$?"; cat gpurun_out/r2c2_probe.log
timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -x -q --tb=short -k "gemm" > gpurun_out/r2c2_pytest_gemm.log 2>&1; echo "pytest gemm exit=$?"; tail -3 gpurun_out/r2c2_pytest_gemm.log
timeout 600 python -m pytest tests/test_model_gpu.py tests/test_ccl_gpu.py -m gpu -x -q --tb=short -s > gpurun_out/r2c2_pytest_model.log 2>&1; echo "pytest model exit=$?"; tail -3 gpurun_out/r2c2_pytest_model.log; grep -E "^e2e|^encoder" gpurun_out/r2c2_pytest_model.log
for cfg in "B200SAM_GEMM_PAIR=1" "B200SAM_GEMM_PAIR=0" "B200SAM_GEMM_PAIR=1 B200SAM_LN_FUSED=0" "B200SAM_GEMM_PAIR=1 B200SAM_ENCODER_OPERANDS=bf16"; do
  tag=$(echo "$cfg" | tr ' =' '__')
  env $cfg timeout 300 python bench.py --steps 10 --warmup 3 --no-refine --no-cpu-baseline > gpurun_out/r2c2_bench_$tag.json 2> gpurun_out/r2c2_bench_$tag.err
  echo "$cfg exit=$? $(python -c "
import json;d=json.load(open('gpurun_out/r2c2_bench_$tag.json'));r=d['roofline']
print(round(d['value'],2), round(d['ms_per_step'],3), d['clocks'].get('sm_mhz'), 'gemmTF', round(r['achieved'],1), {k:v['ms_mean'] for k,v in r['per_shape'].items() if k in ('qkv','proj','lin1','lin2')}, {k:v['ms_mean'] for k,v in r['attention'].items()})" 2>&1)"
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_pair -c 17 -f -o gpurun_out/r2c2_gemm_pair python tools/gemm_pair_probe.py 1 > gpurun_out/r2c2_ncu_pair.log 2>&1; echo "ncu pair exit=$?"

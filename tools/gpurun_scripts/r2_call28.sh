#!/bin/bash
# Round 2, call 28: persistent windowed attention as the default: ops + model tests, bench A/B against the per-window kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_model_gpu.py -m gpu -x -q --tb=short > gpurun_out/r2c28_pytest.log 2>&1; echo "pytest exit=$?"; tail -3 gpurun_out/r2c28_pytest.log
for cfg in "B200SAM_WINATTN=persist" "B200SAM_WINATTN=cta"; do
  tag=$(echo "$cfg" | tr ' =' '__')
  env $cfg timeout 300 python bench.py --steps 10 --warmup 3 --no-refine --no-cpu-baseline > gpurun_out/r2c28_bench_$tag.json 2> gpurun_out/r2c28_bench_$tag.err
  echo "$cfg exit=$? $(python -c "
import json;d=json.load(open('gpurun_out/r2c28_bench_$tag.json'));r=d['roofline']
print(round(d['value'],2), round(d['ms_per_step'],3), d['clocks'].get('sm_mhz'), 'gemmTF', round(r['achieved'],1), {k:v['ms_mean'] for k,v in r['attention'].items()}, d['parity']['dice_min'], d['parity']['embedding_rel_l2'])" 2>&1)"
done

#!/bin/bash
# Round 2, call 40: the direct fp32 epilogue of the pair GEMM actually wired in (+ planes): GEMM probe (modes 0 / 1 / 2), op + model
# tests, encoder bench A/B (B200SAM_GEMM_DIRECT=1 / 0), refine bench, decode launch list
mkdir -p gpurun_out
timeout 300 python tools/gemm_pair_probe.py 20 2>&1 | head -10 | tee gpurun_out/r2c40_probe.log
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_model_gpu.py -m gpu -x -q --tb=short -s > gpurun_out/r2c40_pytest.log 2>&1; echo "pytest exit=$?"; tail -3 gpurun_out/r2c40_pytest.log
grep -h "mismatched\|min dice" gpurun_out/r2c40_pytest.log | head -12
for cfg in "B200SAM_GEMM_DIRECT=1" "B200SAM_GEMM_DIRECT=0" "B200SAM_GEMM_DIRECT=1" "B200SAM_GEMM_DIRECT=0"; do
  tag=$(echo "$cfg" | tr ' =' '__')_$RANDOM
  env $cfg timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2c40_bench_$tag.json 2> gpurun_out/r2c40_bench_$tag.err
  echo "$cfg exit=$? $(python -c "
import json;d=json.load(open('gpurun_out/r2c40_bench_$tag.json'));r=d['roofline']
print(round(d['value'],2), round(d['ms_per_step'],3), d['clocks'].get('sm_mhz'), 'gemmTF', round(r['achieved'],1), {k:v['ms_mean'] for k,v in r['per_shape'].items() if k in ('qkv','proj','lin1','lin2')}, 'refine', round(d['refine']['value']), round(d['refine']['per_image_api']['value']), 'parity', round(d['parity']['dice_min'],5))" 2>&1)"
done
timeout 300 python tools/profile_decode_stage.py 8 all > gpurun_out/r2c40_decode_plain.log 2>&1 && \
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2c40_launches_decode_stage.csv python tools/profile_decode_stage.py 8 all > gpurun_out/r2c40_ncu_decode.log 2>&1
echo "ncu decode exit=$?"

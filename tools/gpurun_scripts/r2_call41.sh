#!/bin/bash
# Round 2, call 41: direct epilogue removed; decoder k | v | q planes GEMM on the CTA-pair kernel (staged epilogue): tests, bench, launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_model_gpu.py -m gpu -x -q --tb=short -s > gpurun_out/r2c41_pytest.log 2>&1; echo "pytest exit=$?"; tail -3 gpurun_out/r2c41_pytest.log
grep -h "mismatched\|min dice" gpurun_out/r2c41_pytest.log | head -12
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2c41_bench.json 2> gpurun_out/r2c41_bench.err
echo "bench exit=$? $(python -c "
import json;d=json.load(open('gpurun_out/r2c41_bench.json'));r=d['roofline']
print(round(d['value'],2), round(d['ms_per_step'],3), 'gemmTF', round(r['achieved'],1), {k:v['ms_mean'] for k,v in r['per_shape'].items() if k in ('qkv','proj','lin1','lin2')}, 'refine', round(d['refine']['value']), round(d['refine']['per_image_api']['value']), 'set500', round(d['set500']['images_per_s'],1), d['set500']['refine_phase'])" 2>&1)"
timeout 300 python tools/profile_decode_stage.py 8 all > gpurun_out/r2c41_decode_plain.log 2>&1 && \
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2c41_launches_decode_stage.csv python tools/profile_decode_stage.py 8 all > gpurun_out/r2c41_ncu_decode.log 2>&1
echo "ncu decode exit=$?"

#!/bin/bash
# Round 2, call 5 (1 GPU): epilogue fixes (early residual loads, unrolled row-stat loads, cheap remote arrive) + npy writer
mkdir -p gpurun_out
timeout 180 python tools/gemm_pair_probe.py 20 > gpurun_out/r2c5_probe.log 2>&1; echo "probe exit=$?"; cat gpurun_out/r2c5_probe.log
timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_model_gpu.py -m gpu -x -q --tb=short > gpurun_out/r2c5_pytest.log 2>&1; echo "pytest exit=$?"; tail -3 gpurun_out/r2c5_pytest.log
for cfg in "B200SAM_LN_FUSED=1" "B200SAM_LN_FUSED=0" "B200SAM_ENCODER_OPERANDS=bf16"; do
  tag=$(echo "$cfg" | tr ' =' '__')
  env $cfg timeout 300 python bench.py --steps 10 --warmup 3 --no-refine --no-cpu-baseline > gpurun_out/r2c5_bench_$tag.json 2> gpurun_out/r2c5_bench_$tag.err
  echo "$cfg exit=$? $(python -c "
import json;d=json.load(open('gpurun_out/r2c5_bench_$tag.json'));r=d['roofline']
print(round(d['value'],2), round(d['ms_per_step'],3), d['clocks'].get('sm_mhz'), 'gemmTF', round(r['achieved'],1), {k:v['ms_mean'] for k,v in r['per_shape'].items() if k in ('qkv','proj','lin1','lin2')}, {k:v['ms_mean'] for k,v in r['attention'].items()})" 2>&1)"
done
timeout 600 python - <<'PY' > gpurun_out/r2c5_pipeline.log 2>&1
import json, torch, sys
sys.path.insert(0, ".")
import bench
dev = torch.device("cuda", 0)
sam = bench.build_model("vit_h", dev)
print(json.dumps(bench.pipeline_throughput(sam, dev)))
PY
echo "pipeline exit=$?"; tail -1 gpurun_out/r2c5_pipeline.log

#!/bin/bash
mkdir -p gpurun_out
for b in 4 6 12 16; do
timeout 600 python bench.py --steps 6 --warmup 3 --batch $b --no-cpu-baseline --no-refine > gpurun_out/bench_b$b.json 2> gpurun_out/bench_b$b.err; python - <<PY
import json
d=json.load(open('gpurun_out/bench_b$b.json'))
print('batch',$b,'value',round(d['value'],1),'e2e', round(d['e2e']['value'],1), d['roofline']['per_shape'], d['clocks']['sm_mhz'])
PY
done
timeout 600 python bench.py --steps 6 --warmup 3 --batch 16 --model vit_l --no-cpu-baseline --no-refine > gpurun_out/bench_vitl.json 2> gpurun_out/bench_vitl.err; python - <<PY
import json
d=json.load(open('gpurun_out/bench_vitl.json'))
print('vit_l b16 value',round(d['value'],1),'e2e', round(d['e2e']['value'],1), d['encoder_frac_of_bf16_peak'], d['roofline']['per_shape'])
PY

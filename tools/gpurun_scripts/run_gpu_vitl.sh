#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --model vit_l --batch 16 --steps 10 --warmup 3 --no-refine --no-cpu-baseline > gpurun_out/bench_vitl.json 2> gpurun_out/bench_vitl.err; echo "exit=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_vitl.json')); print(d['metric'], round(d['value'],1), round(d['e2e']['value'],1), d['ms_per_step'], d['clocks'], d['encoder_frac_of_bf16_peak'])"
timeout 600 python bench.py --model vit_b --batch 16 --steps 10 --warmup 3 --no-refine --no-cpu-baseline > gpurun_out/bench_vitb.json 2> gpurun_out/bench_vitb.err; echo "exit=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_vitb.json')); print(d['metric'], round(d['value'],1), round(d['e2e']['value'],1), d['ms_per_step'], d['encoder_frac_of_bf16_peak'])"

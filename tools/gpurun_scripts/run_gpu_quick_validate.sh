#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit=$?"; tail -1 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench exit=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_final.json'))
print(round(d['value'],1), round(d['e2e']['value'],1), d['clocks'], 'refine', round(d['refine']['value']), 'pipeline', round(d['pipeline']['images_per_s'],1), 'unet', round(d['unet']['images_per_s']))
print({k:(v['frac_of_measured_hbm'], v['ms']) for k,v in d['hbm_stages'].items()})"

#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_ops_gpu.py -m gpu -q -x --tb=short 2>&1 | tail -3
for i in 1 2; do
B200SAM_FORWARD_ONLY=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-refine --no-cpu-baseline > gpurun_out/bench_fwd.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench_fwd.json')); print('forward-only', round(d['value'],2), round(d['e2e']['value'],2), d['clocks']['sm_mhz'], d['roofline']['per_shape'])"
timeout 600 python bench.py --steps 20 --warmup 5 --no-refine --no-cpu-baseline > gpurun_out/bench_bou.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench_bou.json')); print('boustrophedon', round(d['value'],2), round(d['e2e']['value'],2), d['clocks']['sm_mhz'], d['roofline']['per_shape'])"
done

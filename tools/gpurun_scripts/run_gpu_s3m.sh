#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_model_gpu.py -m gpu -q -x --tb=short -k "attention or embedding or encoder or vit" 2>&1 | tail -4
timeout 300 python tools/attn_bench.py 8 2>&1 | grep -E "global \(tcgen05\)|window \(tcgen05\) "
timeout 600 python bench.py --steps 20 --warmup 5 --no-refine --no-cpu-baseline > gpurun_out/bench_poly.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench_poly.json')); print('bench', round(d['value'],2), round(d['e2e']['value'],2), d['clocks']['sm_mhz'])"

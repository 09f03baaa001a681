#!/bin/bash
for b in 4 5 6 8 12; do
timeout 600 python bench.py --steps 20 --warmup 5 --batch $b --no-refine --no-cpu-baseline > gpurun_out/bench_b$b.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench_b$b.json')); print('batch $b', round(d['value'],2), round(d['e2e']['value'],2), d['clocks']['sm_mhz'])"
done

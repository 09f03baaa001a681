#!/bin/bash
# Round 2, call 29: persistent vs per-window windowed attention, alternating bench runs (step time is what counts: the persistent
# kernel gives up the tail overlap with the next kernel's programmatic dependent launch)
mkdir -p gpurun_out
for rep in 1 2 3; do
for cfg in "B200SAM_WINATTN=persist" "B200SAM_WINATTN=cta"; do
  tag=$(echo "$cfg" | tr ' =' '__')_$rep
  env $cfg timeout 300 python bench.py --steps 20 --warmup 5 --no-refine --no-cpu-baseline > gpurun_out/r2c29_bench_$tag.json 2> gpurun_out/r2c29_bench_$tag.err
  echo "$cfg rep $rep exit=$? $(python -c "
import json;d=json.load(open('gpurun_out/r2c29_bench_$tag.json'));r=d['roofline']
print(round(d['value'],2), round(d['ms_per_step'],3), d['clocks'].get('sm_mhz'), 'gemmTF', round(r['achieved'],1), {k:v['ms_mean'] for k,v in r['attention'].items()})" 2>&1)"
done
done

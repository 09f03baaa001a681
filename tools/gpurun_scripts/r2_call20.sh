#!/bin/bash
# Round 2, call 20: staged pipeline with persistent streams; pipeline + writer with the GIL held across library calls vs released
mkdir -p gpurun_out
timeout 600 python tools/overlap_probe.py 192 2>&1 | tee gpurun_out/r2c20_overlap.log
for cfg in "B200SAM_RELEASE_GIL=0" "B200SAM_RELEASE_GIL=1"; do
  tag=$(echo "$cfg" | tr ' =' '__')
  env $cfg timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2c20_bench_$tag.json 2> gpurun_out/r2c20_bench_$tag.err
  echo "$cfg exit=$? $(python -c "
import json;d=json.load(open('gpurun_out/r2c20_bench_$tag.json'))
print(round(d['value'],2), 'latency_b1', d['latency_b1']['ms_median'], 'refine', round(d['refine']['value']), d['refine']['per_image_api']['value'], 'pipeline', d['pipeline']['images_per_s'], d['pipeline']['ms_per_image_samples'], d['pipeline']['with_async_writer'])" 2>&1)"
done

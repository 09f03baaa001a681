#!/bin/bash
# Round 2, call 46: A/B of the upload rings (pinned + copy stream) against pageable uploads on the compute stream
mkdir -p gpurun_out
for cfg in "B200SAM_UPLOAD_RINGS=1" "B200SAM_UPLOAD_RINGS=0" "B200SAM_UPLOAD_RINGS=1" "B200SAM_UPLOAD_RINGS=0"; do
  tag=$(echo "$cfg" | tr ' =' '__')_$RANDOM
  env $cfg timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2c46_bench_$tag.json 2> gpurun_out/r2c46_bench_$tag.err
  echo "$cfg exit=$? $(python -c "
import json;d=json.load(open('gpurun_out/r2c46_bench_$tag.json'))
print(round(d['value'],1), 'set500', round(d['set500']['images_per_s'],1), round(d['set500']['embed_phase']['embeds_per_s'],1), 'pipeline', round(d['pipeline']['images_per_s'],1), 'native', round(d['pipeline']['embed_native_2570x2040']['images_per_s'],1), 'writer', round(d['pipeline']['with_async_writer']['images_per_s'],1))" 2>&1)"
done

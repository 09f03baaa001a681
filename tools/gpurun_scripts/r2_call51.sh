#!/bin/bash
# Round 2, call 51: native-resolution embedding leg over 64 images (pipeline fill amortised)
mkdir -p gpurun_out
for rep in 1 2; do
timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2c51_bench_$rep.json 2> gpurun_out/r2c51_bench_$rep.err
echo "bench exit=$? $(python -c "
import json;d=json.load(open('gpurun_out/r2c51_bench_$rep.json'))
print(round(d['value'],1), 'set500', round(d['set500']['images_per_s'],1), round(d['set500']['embed_phase']['embeds_per_s'],1), 'pipeline', round(d['pipeline']['images_per_s'],1), 'native', round(d['pipeline']['embed_native_2570x2040']['images_per_s'],1), 'writer', round(d['pipeline']['with_async_writer']['images_per_s'],1), 'refine_phase', d['set500']['refine_phase']['masks_per_s'])" 2>&1)"
done

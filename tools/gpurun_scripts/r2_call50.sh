#!/bin/bash
# Round 2, call 50: pipeline leg hands the masks to the host through the in-memory writer (pinned D2H on a side stream) instead of .cpu() per image
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_storage.py -m gpu -x -q --tb=short > gpurun_out/r2c50_pytest.log 2>&1; echo "pytest exit=$?"; tail -3 gpurun_out/r2c50_pytest.log
for rep in 1 2; do
timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2c50_bench_$rep.json 2> gpurun_out/r2c50_bench_$rep.err
echo "bench exit=$? $(python -c "
import json;d=json.load(open('gpurun_out/r2c50_bench_$rep.json'))
print(round(d['value'],1), 'set500', round(d['set500']['images_per_s'],1), round(d['set500']['embed_phase']['embeds_per_s'],1), 'pipeline', round(d['pipeline']['images_per_s'],1), 'native', round(d['pipeline']['embed_native_2570x2040']['images_per_s'],1), 'writer', round(d['pipeline']['with_async_writer']['images_per_s'],1), 'refine_phase', d['set500']['refine_phase']['masks_per_s'])" 2>&1)"
done

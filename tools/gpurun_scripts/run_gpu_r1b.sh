#!/bin/bash
# iteration: full GPU test-suite, HBM-stage bench, short bench (no CPU baseline)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit=$?"; tail -15 gpurun_out/pytest_gpu.log
timeout 300 python tools/stage_bench.py > gpurun_out/stage.json 2> gpurun_out/stage.err; echo "stage exit=$?"; cat gpurun_out/stage.json | head -40
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_iter.json 2> gpurun_out/bench_iter.err; echo "bench exit=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_iter.json'))
print('value',d['value'],'e2e', d['e2e']['value'], d['roofline']['per_shape'], 'refine',d['refine'], d['clocks'])
PY
tail -3 gpurun_out/bench_iter.err

#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -q --tb=short -x -k "gemm" > gpurun_out/pytest_gemm.log 2>&1; echo "pytest exit=$?"; tail -4 gpurun_out/pytest_gemm.log | cut -c1-300
timeout 900 python -m pytest tests/test_model_gpu.py -m gpu -q --tb=short -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit=$?"; tail -4 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 900 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_iter.json 2> gpurun_out/bench_iter.err; echo "bench exit=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_iter.json'))
print('value',d['value'],'e2e', d['e2e']['value'], d['roofline']['per_shape'], 'refine',d['refine']['value'], d['refine']['per_image_api']['value'], d['clocks'], d.get('pipeline'), d.get('hbm_stages'))
PY
tail -3 gpurun_out/bench_iter.err

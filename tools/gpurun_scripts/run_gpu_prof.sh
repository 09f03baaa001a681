#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -q -k "gemm" --tb=short > gpurun_out/ops_gemm.log 2>&1; echo "gemm tests exit=$?"; tail -3 gpurun_out/ops_gemm.log
timeout 900 python bench.py --steps 5 --warmup 3 --batch 8 --no-cpu-baseline > gpurun_out/bench_b8.json 2> gpurun_out/bench_b8.err; echo "bench exit=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_b8.json'))
print(d['value'], d['e2e']['value'], d['roofline']['per_shape'], d['refine']['value'], d['clocks'])
PY
timeout 600 python tools/profile_encoder.py --batch 8 --iters 2 > gpurun_out/prof_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 232 -c 232 --csv --log-file gpurun_out/launches_enc_b8.csv python tools/profile_encoder.py --batch 8 --iters 2 > gpurun_out/ncu_enc.log 2>&1
echo "ncu exit=$?"; tail -2 gpurun_out/ncu_enc.log

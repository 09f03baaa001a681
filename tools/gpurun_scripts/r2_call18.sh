#!/bin/bash
# Round 2, call 18: full GPU suite + full bench (global attention with P through TMEM, writer with size-class pinned pool,
# overlapped embed + refine leg of set500)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --tb=short > gpurun_out/r2c18_pytest.log 2>&1; echo "pytest exit=$?"; tail -3 gpurun_out/r2c18_pytest.log
timeout 900 python bench.py > gpurun_out/r2c18_bench.json 2> gpurun_out/r2c18_bench.err; echo "bench exit=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2c18_bench.json'))
for k in ('value','ms_per_step','dtype','clocks','e2e','parity','set500','latency_b1','vit_l_batch16','refine','pipeline','hbm_stages','cpu_baseline'):
    print(k, d.get(k))
r=d['roofline']; print({k:v for k,v in r.items() if k not in ('per_shape','attention')}); print(r.get('per_shape')); print(r.get('attention'))
PY

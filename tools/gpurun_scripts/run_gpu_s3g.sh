#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench exit=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_final.json'))
print({k:d[k] for k in ['value','ms_per_step','clocks']})
print('e2e',d['e2e']['value'],'roofline',d['roofline']['frac'],d['roofline']['per_shape'])
print('refine',d['refine']['value'],'pipeline',d['pipeline']['images_per_s'],'unet',d['unet']['images_per_s'])
print(d['hbm_stages'])
PY
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench default exit=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_default.json')); print(d['value'], d['steps'], d['warmup'], d['e2e']['value'], d['clocks'])"

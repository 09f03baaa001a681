#!/bin/bash
# Round 2, call 11: whole-row windowed attention kernel (2 threads per row) vs the 64-key-tile kernel
mkdir -p gpurun_out
B200SAM_WINATTN=tiles timeout 120 python tools/attention_probe.py 8 fp16 save gpurun_out/r2c11_att.pt 2>&1 | tee gpurun_out/r2c11_probe_tiles.log
timeout 120 python tools/attention_probe.py 8 fp16 check gpurun_out/r2c11_att.pt 2>&1 | tee gpurun_out/r2c11_probe_rows.log
rm -f gpurun_out/r2c11_att.pt
timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -x -q --tb=short -k "attention" > gpurun_out/r2c11_pytest_att.log 2>&1; echo "pytest attention exit=$?"; tail -5 gpurun_out/r2c11_pytest_att.log
timeout 600 python -m pytest tests/test_model_gpu.py -m gpu -x -q --tb=short > gpurun_out/r2c11_pytest_model.log 2>&1; echo "pytest model exit=$?"; tail -3 gpurun_out/r2c11_pytest_model.log
for cfg in "B200SAM_WINATTN=rows" "B200SAM_WINATTN=tiles"; do
  tag=$(echo "$cfg" | tr ' =' '__')
  env $cfg timeout 300 python bench.py --steps 10 --warmup 3 --no-refine --no-cpu-baseline > gpurun_out/r2c11_bench_$tag.json 2> gpurun_out/r2c11_bench_$tag.err
  echo "$cfg exit=$? $(python -c "
import json;d=json.load(open('gpurun_out/r2c11_bench_$tag.json'));r=d['roofline']
print(round(d['value'],2), round(d['ms_per_step'],3), d['clocks'].get('sm_mhz'), 'gemmTF', round(r['achieved'],1), {k:v['ms_mean'] for k,v in r['attention'].items()})" 2>&1)"
done

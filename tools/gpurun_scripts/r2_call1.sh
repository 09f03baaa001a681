#!/bin/bash
# Round 2, call 1: validate fp16 operands + LN folding + PDL; e2e Dice; A/B bench numbers.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r2c1_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q --tb=short -s > gpurun_out/r2c1_pytest.log 2>&1; echo "pytest exit=$?"
grep -E "passed|failed|error" gpurun_out/r2c1_pytest.log | tail -3
grep -E "^e2e|^encoder|^attention|^refine|^predict" gpurun_out/r2c1_pytest.log | head -60
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2c1_smoke.log 2>&1; echo "smoke exit=$?"; tail -2 gpurun_out/r2c1_smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2c1_bench_default.json 2> gpurun_out/r2c1_bench_default.err; echo "bench default exit=$?"
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2c1_bench_default.json"))
    print({k: d[k] for k in ("value", "ms_per_step", "dtype", "clocks")})
    print("e2e", d["e2e"]["value"], "parity", d.get("parity"))
    r = d["roofline"]; print({k: r[k] for k in ("achieved", "frac", "frac_of_burst_peak", "share_of_step", "instrumented_ms_per_step")}); print(r["per_shape"]); print(r["attention"])
except Exception as e:
    print("parse failed", e)
PY
for cfg in "B200SAM_PDL=0" "B200SAM_LN_FUSED=0" "B200SAM_ENCODER_OPERANDS=bf16" "B200SAM_PDL=1"; do
  env $cfg timeout 300 python bench.py --steps 10 --warmup 3 --no-refine --no-cpu-baseline > gpurun_out/r2c1_bench_$cfg.json 2> gpurun_out/r2c1_bench_$cfg.err
  echo "$cfg exit=$? $(python -c "import json;d=json.load(open('gpurun_out/r2c1_bench_$cfg.json'));print(round(d['value'],2), round(d['ms_per_step'],3), d['clocks'].get('sm_mhz'), round(d['roofline']['achieved'],1))" 2>&1)"
done

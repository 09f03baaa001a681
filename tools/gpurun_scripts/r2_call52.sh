#!/bin/bash
# Round 2, call 52 (1 GPU): state after the attention work: full GPU tests, smoke, full default bench, encoder launch list,
# ncu --set full of the two attention kernels (global with P through TMEM and 8x8 key blocks)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --tb=short -s > gpurun_out/r2c52_pytest.log 2>&1; echo "pytest exit=$?"; grep -E "passed|failed|error" gpurun_out/r2c52_pytest.log | tail -3; grep -E "^e2e|^medsam|^encoder vit_h" gpurun_out/r2c52_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2c52_smoke.log 2>&1; echo "smoke exit=$?"; tail -1 gpurun_out/r2c52_smoke.log
timeout 900 python bench.py > gpurun_out/r2c52_bench.json 2> gpurun_out/r2c52_bench.err; echo "bench exit=$?"; tail -3 gpurun_out/r2c52_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2c52_bench.json"))
for k in ("value", "ms_per_step", "dtype", "clocks", "e2e", "parity", "set500", "latency_b1", "vit_l_batch16", "refine", "pipeline", "hbm_stages", "cpu_baseline", "gpu_launches"):
    print(k, d.get(k))
r = d["roofline"]; print({k: r[k] for k in r if k not in ("per_shape", "attention", "how", "kernel")}); print(r["per_shape"]); print(r["attention"])
PY


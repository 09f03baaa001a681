#!/bin/bash
# Round 2, call 43: final default bench run (all legs) on the final tree
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r2c43_bench.json 2> gpurun_out/r2c43_bench.err; echo "bench exit=$?"; tail -2 gpurun_out/r2c43_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2c43_bench.json"))
for k in ("value", "ms_per_step", "dtype", "clocks", "e2e", "parity", "set500", "latency_b1", "vit_l_batch16", "refine", "pipeline", "hbm_stages", "cpu_baseline", "gpu_launches"):
    print(k, d.get(k))
r = d["roofline"]; print({k: r[k] for k in r if k not in ("per_shape", "attention", "how")}); print(r["per_shape"]); print(r["attention"])
PY

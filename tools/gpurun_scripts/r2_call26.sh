#!/bin/bash
# Round 2, call 26 (1 GPU): state after the attention work: full GPU tests, smoke, full default bench, encoder launch list,
# ncu --set full of the two attention kernels (global with P through TMEM and 8x8 key blocks)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --tb=short -s > gpurun_out/r2c26_pytest.log 2>&1; echo "pytest exit=$?"; grep -E "passed|failed|error" gpurun_out/r2c26_pytest.log | tail -3; grep -E "^e2e|^medsam|^encoder vit_h" gpurun_out/r2c26_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2c26_smoke.log 2>&1; echo "smoke exit=$?"; tail -1 gpurun_out/r2c26_smoke.log
timeout 900 python bench.py > gpurun_out/r2c26_bench.json 2> gpurun_out/r2c26_bench.err; echo "bench exit=$?"; tail -3 gpurun_out/r2c26_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2c26_bench.json"))
for k in ("value", "ms_per_step", "dtype", "clocks", "e2e", "parity", "set500", "latency_b1", "vit_l_batch16", "refine", "pipeline", "hbm_stages", "cpu_baseline", "gpu_launches"):
    print(k, d.get(k))
r = d["roofline"]; print({k: r[k] for k in r if k not in ("per_shape", "attention", "how", "kernel")}); print(r["per_shape"]); print(r["attention"])
PY
timeout 300 python tools/profile_encoder.py --batch 8 --iters 2 > gpurun_out/r2c26_prof_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:b200sam -s 167 -c 167 --csv --log-file gpurun_out/r2c26_launches_enc_b8.csv python tools/profile_encoder.py --batch 8 --iters 2 > gpurun_out/r2c26_ncu_enc.log 2>&1
echo "ncu launches exit=$?"
timeout 120 python tools/profile_attention.py 8 fp16 > gpurun_out/r2c26_attn_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_tc -s 2 -c 2 -f -o gpurun_out/r2c26_attn python tools/profile_attention.py 8 fp16 > gpurun_out/r2c26_ncu_attn.log 2>&1
echo "ncu attn exit=$?"

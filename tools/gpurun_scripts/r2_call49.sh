#!/bin/bash
# Round 2, call 49 (2 GPUs): full GPU test suite on rank-0 GPU, then the 2-GPU bench contract run (weak headline + set500 with gathers)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --tb=short > gpurun_out/r2c49_pytest.log 2>&1; echo "pytest exit=$?"; tail -2 gpurun_out/r2c49_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2c49_smoke.log 2>&1; echo "smoke exit=$?"; tail -1 gpurun_out/r2c49_smoke.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2c49_bench_2gpu.json 2> gpurun_out/r2c49_bench_2gpu.err; echo "bench 2gpu exit=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2c49_bench_2gpu.json"))
for k in ("value", "n_gpus", "ms_per_step", "e2e", "set500"):
    print(k, d.get(k))
PY

#!/bin/bash
# 8-GPU check of the bench contract (weak-scaling headline + the 500-image strong-scaling set with both gathers)
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2_bench_8gpu.json 2> gpurun_out/r2_bench_8gpu.err; echo "bench 8gpu exit=$?"
tail -3 gpurun_out/r2_bench_8gpu.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_bench_8gpu.json"))
for k in ("value", "n_gpus", "ms_per_step", "e2e", "set500", "clocks"):
    print(k, d.get(k))
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 8 --steps 2 --warmup 1 > gpurun_out/r2_bench_8gpu_ref.json 2> gpurun_out/r2_bench_8gpu_ref.err; echo "reference arm exit=$?"; cat gpurun_out/r2_bench_8gpu_ref.json | cut -c1-400

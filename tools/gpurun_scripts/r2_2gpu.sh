#!/bin/bash
# 2-GPU check of the bench contract (weak-scaling headline + the 500-image strong-scaling set with both gathers)
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/r2_bench_2gpu.err; echo "bench 2gpu exit=$?"
tail -3 gpurun_out/r2_bench_2gpu.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_bench_2gpu.json"))
for k in ("value", "n_gpus", "ms_per_step", "e2e", "set500", "clocks"):
    print(k, d.get(k))
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2_bench_2gpu_ref.json 2> gpurun_out/r2_bench_2gpu_ref.err; echo "reference arm exit=$?"; cat gpurun_out/r2_bench_2gpu_ref.json | cut -c1-400
timeout 300 python -m pytest tests/test_model_gpu.py -m gpu -q -k "second_device" > gpurun_out/r2_2gpu_pytest.log 2>&1; echo "second-device test exit=$?"; tail -2 gpurun_out/r2_2gpu_pytest.log

#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "bench2 exit=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_2gpu.json')); print(d['n_gpus'], d['value'], d['e2e']['value'], d['clocks'], d['config']['parallelism'])"
tail -3 gpurun_out/bench_2gpu.err

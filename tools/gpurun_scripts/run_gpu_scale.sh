#!/bin/bash
# 1 -> 8 GPU weak scaling of the headline bench on one box (run with `gpurun --gpus 8`).
mkdir -p gpurun_out
for n in 8 4 2 1; do
  if [ $n -eq 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-refine --no-cpu-baseline > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n --steps 10 --warmup 3 --no-refine --no-cpu-baseline > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err
  fi
  echo "n=$n exit=$?"
  python -c "
import json; d=json.load(open('gpurun_out/scale_$n.json')); print(d['n_gpus'], round(d['value'],1), round(d['e2e']['value'],1), d['ms_per_step'], d['clocks'].get('sm_mhz'))"
done

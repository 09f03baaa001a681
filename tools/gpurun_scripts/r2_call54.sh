#!/bin/bash
# Round 2, call 54: stress of the persistent windowed attention kernel and the encoder loop (rare races / hangs): the attention
# op tests 30 times, 200 encoder steps, the model tests 3 times
mkdir -p gpurun_out
fail=0
for i in $(seq 1 30); do
  timeout 120 python -m pytest tests/test_ops_gpu.py -m gpu -x -q --tb=line -k "attention" > gpurun_out/r2c54_att_$i.log 2>&1 || { fail=$((fail+1)); echo "attention run $i FAILED"; tail -3 gpurun_out/r2c54_att_$i.log; }
done
echo "attention stress: $fail failures of 30"
rm -f gpurun_out/r2c54_att_*.log
timeout 600 python bench.py --steps 200 --warmup 5 --no-cpu-baseline --no-refine > gpurun_out/r2c54_bench200.json 2> gpurun_out/r2c54_bench200.err; echo "bench 200 steps exit=$?"
python -c "
import json;d=json.load(open('gpurun_out/r2c54_bench200.json'));print(round(d['value'],2), d['ms_per_step'], d['clocks'])"
for i in 1 2 3; do timeout 600 python -m pytest tests/test_model_gpu.py -m gpu -x -q --tb=line > gpurun_out/r2c54_model_$i.log 2>&1; echo "model run $i exit=$?"; tail -1 gpurun_out/r2c54_model_$i.log; done

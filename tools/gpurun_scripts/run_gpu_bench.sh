#!/bin/bash
mkdir -p gpurun_out
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit=$?" >> gpurun_out/smoke.log; tail -3 gpurun_out/smoke.log
for B in 1 4 8; do
  timeout 900 python bench.py --steps 5 --warmup 3 --batch $B --no-cpu-baseline > gpurun_out/bench_b$B.json 2> gpurun_out/bench_b$B.err; echo "bench B=$B exit=$?"; tail -c 3000 gpurun_out/bench_b$B.json; tail -5 gpurun_out/bench_b$B.err
done

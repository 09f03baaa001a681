#!/bin/bash
# Full GPU validation + the profile artefacts that go under profiles/ (one gpurun call).
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit=$?"; tail -1 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench exit=$?"; head -c 5000 gpurun_out/bench_final.json; echo
timeout 600 python tools/profile_encoder.py --batch 8 --iters 2 > gpurun_out/prof_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 232 -c 232 --csv --log-file gpurun_out/launches_enc_b8.csv python tools/profile_encoder.py --batch 8 --iters 2 > gpurun_out/ncu_enc.log 2>&1
echo "ncu launches exit=$?"
timeout 300 python tools/profile_unet.py 8 > gpurun_out/pu_plain.log 2>&1 && \
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_unet.csv python tools/profile_unet.py 8 > gpurun_out/ncu_unet.log 2>&1
echo "ncu unet exit=$?"
timeout 300 python tools/profile_gemm.py > gpurun_out/pg_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 4 -c 4 -f -o gpurun_out/gemm_final python tools/profile_gemm.py > gpurun_out/ncu_gemm.log 2>&1
echo "ncu gemm exit=$?"
timeout 600 python tools/timeline_encoder.py 8 > gpurun_out/timeline_enc.log 2>&1; echo "timeline exit=$?"; grep -A8 "^kernels" gpurun_out/timeline_enc.log

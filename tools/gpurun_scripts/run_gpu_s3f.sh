#!/bin/bash
# Round-1 (session 3) validation + profile artefacts
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit=$?"; tail -1 gpurun_out/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench exit=$?"; head -c 4500 gpurun_out/bench_final.json; echo
timeout 600 python tools/profile_encoder.py --batch 8 --iters 2 > gpurun_out/prof_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 232 -c 232 --csv --log-file gpurun_out/launches_enc_b8.csv python tools/profile_encoder.py --batch 8 --iters 2 > gpurun_out/ncu_enc.log 2>&1
echo "ncu launches exit=$?"
timeout 300 python tools/profile_decode_stage.py 8 all > gpurun_out/pds_plain.log 2>&1 && \
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_decode_stage.csv python tools/profile_decode_stage.py 8 all > gpurun_out/ncu_pds.log 2>&1
echo "ncu decode launches exit=$?"
timeout 300 python tools/profile_decode_stage.py 8 stages > gpurun_out/pds2_plain.log 2>&1 && \
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:'upscale_mask_fast|prompt_accum|resize_' -f -o gpurun_out/stages_final python tools/profile_decode_stage.py 8 stages > gpurun_out/ncu_pds2.log 2>&1
echo "ncu stages exit=$?"

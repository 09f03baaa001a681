#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:'prompt_accum|prompt_init|prompt_finalize' --csv --log-file gpurun_out/pe_ncu.csv python tools/profile_decode_stage.py 8 stages > gpurun_out/ncu_pe.log 2>&1; echo "exit=$?"
grep -v "^==" gpurun_out/pe_ncu.csv | cut -d, -f5,13,15 | tail -12

#!/bin/bash
# Round 2, call 13: whole-row windowed attention, second iteration (shared-memory exchanges, 32-column TMEM loads, FMA-pipe exp2)
mkdir -p gpurun_out
B200SAM_WINATTN=tiles timeout 120 python tools/attention_probe.py 8 fp16 save gpurun_out/r2c13_att.pt 2>&1 | tee gpurun_out/r2c13_probe_tiles.log
timeout 120 python tools/attention_probe.py 8 fp16 check gpurun_out/r2c13_att.pt 2>&1 | tee gpurun_out/r2c13_probe_rows.log
rm -f gpurun_out/r2c13_att.pt
timeout 120 python tools/window_attention_timeline.py 8 2>&1 | tee gpurun_out/r2c13_timeline.log
timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -x -q --tb=short -k "attention" > gpurun_out/r2c13_pytest_att.log 2>&1; echo "pytest attention exit=$?"; tail -3 gpurun_out/r2c13_pytest_att.log

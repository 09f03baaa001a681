#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/timeline_encoder.py 8 > gpurun_out/timeline_enc.log 2>&1; echo "exit=$?"; tail -16 gpurun_out/timeline_enc.log
timeout 300 python tools/stage_bench.py > gpurun_out/stage.json 2> gpurun_out/stage.err; echo "stage exit=$?"; grep -A6 ingest gpurun_out/stage.json

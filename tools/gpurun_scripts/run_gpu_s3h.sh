#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/pipeline_phases.py > gpurun_out/pipeline_phases.log 2>&1; echo "exit=$?"; head -60 gpurun_out/pipeline_phases.log | cut -c1-180

#!/bin/bash
# Round 2, call 19: staged embed + refine pipeline: stage size sweep, with / without the second stream
mkdir -p gpurun_out
timeout 600 python tools/overlap_probe.py 192 2>&1 | tee gpurun_out/r2c19_overlap.log

#!/bin/bash
timeout 600 python -m pytest tests/test_resize_gpu.py -m gpu -q -x --tb=short 2>&1 | tail -2
timeout 300 python tools/stage_bench.py 2>/dev/null | grep -A6 ingest

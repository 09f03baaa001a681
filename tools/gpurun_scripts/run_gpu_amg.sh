#!/bin/bash
timeout 600 python -m pytest tests/test_amg_gpu.py -m gpu -q -x --tb=short 2>&1 | tail -8

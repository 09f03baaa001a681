#!/bin/bash
timeout 900 python -m pytest tests/test_unet_gpu.py -m gpu -q -x --tb=short -s 2>&1 | tail -15

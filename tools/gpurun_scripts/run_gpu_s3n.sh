#!/bin/bash
timeout 300 python tools/attn_bench.py 8 2>&1 | grep -E "global \(tcgen05\)"

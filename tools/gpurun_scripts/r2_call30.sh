#!/bin/bash
# Round 2, call 30: the encoder's four GEMM shapes: our kernels (with their epilogues) vs the library GEMM (torch.matmul, fp16)
mkdir -p gpurun_out
timeout 300 python tools/gemm_pair_probe.py 20 2>&1 | tee gpurun_out/r2c30_probe.log

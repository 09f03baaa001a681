#!/bin/bash
# Round 2, call 17: global attention variants (P through TMEM): one thread per row vs two threads per row with pair-wise barriers;
# overlapped embed + refine pipeline test
mkdir -p gpurun_out
timeout 120 python tools/attention_probe.py 8 fp16 save gpurun_out/r2c17_att.pt 2>&1 | tee gpurun_out/r2c17_probe_row.log
B200SAM_GLOBATTN=pair timeout 120 python tools/attention_probe.py 8 fp16 check gpurun_out/r2c17_att.pt 2>&1 | tee gpurun_out/r2c17_probe_pair.log
rm -f gpurun_out/r2c17_att.pt
B200SAM_GLOBATTN=pair timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -x -q --tb=short -k "attention" > gpurun_out/r2c17_pytest_att.log 2>&1; echo "pytest attention (pair) exit=$?"; tail -3 gpurun_out/r2c17_pytest_att.log
timeout 600 python -m pytest tests/test_model_gpu.py -m gpu -x -q --tb=short -k "overlapped or pipeline" > gpurun_out/r2c17_pytest_pipe.log 2>&1; echo "pytest pipeline exit=$?"; tail -5 gpurun_out/r2c17_pytest_pipe.log

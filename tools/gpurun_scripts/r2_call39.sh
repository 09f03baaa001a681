#!/bin/bash
# Round 2, call 39: decoder N >= 256 GEMMs (i2t out projection, token MLP) on the CTA-pair kernel (a_wrap support): tests, refine bench, launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_ops_gpu.py -m gpu -x -q --tb=short -s > gpurun_out/r2c39_pytest.log 2>&1; echo "pytest exit=$?"; tail -3 gpurun_out/r2c39_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2c39_bench.json 2> gpurun_out/r2c39_bench.err; echo "bench exit=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2c39_bench.json"))
print("value", d["value"]); print("refine", d["refine"]["value"], d["refine"]["ms_per_image"], d["refine"]["per_image_api"]); print("set500", d["set500"]["images_per_s"], d["set500"]["refine_phase"]); print("pipeline", d["pipeline"]["images_per_s"]); 
PY
timeout 300 python tools/profile_decode_stage.py 8 all > gpurun_out/r2c39_decode_plain.log 2>&1 && \
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2c39_launches_decode_stage.csv python tools/profile_decode_stage.py 8 all > gpurun_out/r2c39_ncu_decode.log 2>&1
echo "ncu decode exit=$?"
grep -h "mismatched\|min dice\|Dice" gpurun_out/r2c39_pytest.log | head -12

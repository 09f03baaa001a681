#!/bin/bash
# Round 2, call 45: native-resolution upload ring (pinned + copy stream): pipeline tests, bench pipeline leg incl. embed_native
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_resize_gpu.py -m gpu -x -q --tb=short > gpurun_out/r2c45_pytest.log 2>&1; echo "pytest exit=$?"; tail -3 gpurun_out/r2c45_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2c45_bench.json 2> gpurun_out/r2c45_bench.err; echo "bench exit=$?"; tail -2 gpurun_out/r2c45_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2c45_bench.json"))
print("value", d["value"]); print("set500", d["set500"]["images_per_s"]); print("pipeline", d["pipeline"])
PY

#!/bin/bash
# Round 2, call 44: double-buffered pinned upload of host image batches in generate_img_embeddings: pipeline tests, bench (set500, pipeline)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_model_gpu.py -m gpu -x -q --tb=short -k "pipeline or overlapped or e2e" > gpurun_out/r2c44_pytest.log 2>&1; echo "pytest exit=$?"; tail -3 gpurun_out/r2c44_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2c44_bench.json 2> gpurun_out/r2c44_bench.err; echo "bench exit=$?"; tail -2 gpurun_out/r2c44_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2c44_bench.json"))
print("value", d["value"], "e2e", d["e2e"]["value"]); print("set500", d["set500"]["images_per_s"], d["set500"]["embed_phase"], d["set500"]["refine_phase"]); print("pipeline", d["pipeline"])
PY

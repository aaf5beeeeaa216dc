#!/bin/bash
# r2x: the whole GPU suite on the build with the device SweepBuilder, then the default bench line (tree_build record included)
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2x_pytest_gpu.log 2>&1
tail -5 gpurun_out/r2x_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/r2x_bench_default_1gpu.json 2> gpurun_out/r2x_bench_default_1gpu.err
python - <<'PY'
import json
line = json.load(open("gpurun_out/r2x_bench_default_1gpu.json"))
print("value", line["value"], "frac", line["roofline"]["frac"], "e2e", line["e2e"]["value"])
print("tree_build", json.dumps(line.get("tree_build")))
for key, record in line.get("render", {}).items():
    if isinstance(record, dict):
        print(key, record["value"] / 1e6, record["roofline"]["frac"], record["e2e"]["value"] / 1e6, record["cpu_baseline"]["value"] / 1e6)
PY

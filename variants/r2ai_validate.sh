#!/bin/bash
# r2ai (r2ae again after ECHO_POOL 256 -> 128): the last validation of round 2 — the whole GPU suite, the default bench line exactly as the driver launches it (light_tree_build
# record included), then the ncu launch list of the C2 command on the same build (after its plain run exited 0)
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2ai_pytest_gpu.log 2>&1
tail -5 gpurun_out/r2ai_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/r2ai_bench_default_1gpu.json 2> gpurun_out/r2ai_bench_default_1gpu.err
python - <<'PY'
import json
line = json.load(open("gpurun_out/r2ai_bench_default_1gpu.json"))
print("value", line["value"], "frac", line["roofline"]["frac"], "e2e", line["e2e"]["value"], "cpu", line["cpu_baseline"]["value"])
print("tree_build", json.dumps(line.get("tree_build")))
for key, record in line.get("render", {}).items():
    if isinstance(record, dict):
        print(key, record["value"] / 1e6, record["roofline"]["frac"], record["e2e"]["value"] / 1e6, record["cpu_baseline"]["value"] / 1e6)
        if "light_tree_build" in record:
            print("light_tree_build", json.dumps(record["light_tree_build"]))
PY
TRACE="python bench.py --workload trace --steps 2 --warmup 3 --no-cpu-baseline --no-secondary --no-tree-build"
$TRACE > gpurun_out/r2ai_trace_plain.json 2> gpurun_out/r2ai_trace_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r2ai_launches_trace.csv $TRACE > gpurun_out/r2ai_trace_ncu.log 2>&1
tail -3 gpurun_out/r2ai_trace_ncu.log
# DRAM traffic per launch of the two C2 kernels on THIS build (profiles/ncu_traffic.json is what bench.py's roofline.traffic reads)
$TRACE > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:persistent_batch_kernel -s 2 -c 2 -f -o gpurun_out/r2ai_prof_trace $TRACE > gpurun_out/r2ai_trace_ncu2.log 2>&1
tail -2 gpurun_out/r2ai_trace_ncu2.log; ls -la gpurun_out/r2ai_prof_trace.ncu-rep

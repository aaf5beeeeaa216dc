#!/bin/bash
# r2w: the device SweepBuilder (sweep.cu): GPU build tests, timing breakdown at C2 / C5 size, C2 headline on the device-built tree
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_build.py -m gpu -x -q -s > gpurun_out/r2w_pytest_build.log 2>&1
tail -5 gpurun_out/r2w_pytest_build.log
ECHO_B200_PROFILE=1 timeout 600 python - > gpurun_out/r2w_build_times.log 2>&1 <<'PY'
import time, numpy as np
from echorenderer_b200 import build_qbvh_device, host, scenes, _native
for name, description in (("c2", scenes.terrain_scene()), ("c5", scenes.large_scene())):
    build_qbvh_device(description.triangles[:64], description.spheres[:0])
    for algorithm in (2, 2, 1, 0):
        _native.set_option("BUILD_ALGORITHM", algorithm)
        started = time.perf_counter()
        nodes, depth = build_qbvh_device(description.triangles, description.spheres)
        print(name, "algorithm", algorithm, "call %.1f ms" % ((time.perf_counter() - started) * 1e3), len(nodes), depth, flush=True)
    _native.set_option("BUILD_ALGORITHM", 2)
    started = time.perf_counter()
    expected, expected_depth = host.build_qbvh(description.triangles, description.spheres)
    print(name, "host mirror %.1f ms" % ((time.perf_counter() - started) * 1e3), len(expected), expected_depth, flush=True)
    nodes, depth = build_qbvh_device(description.triangles, description.spheres)
    print(name, "identical", nodes.tobytes() == expected.tobytes() and depth == expected_depth, flush=True)
PY
cat gpurun_out/r2w_build_times.log | grep -v "^\[echo_b200 build\] [0-9]* prim" | tail -30
timeout 600 python bench.py --workload trace --tree device --steps 10 --no-cpu-baseline > gpurun_out/r2w_trace_device_tree.json 2> gpurun_out/r2w_trace_device_tree.err
tail -c 600 gpurun_out/r2w_trace_device_tree.json

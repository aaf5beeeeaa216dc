#!/bin/bash
# r2z: device SweepBuilder with warp-aggregated cost atomics: build tests + timing
set -x
timeout 900 python -m pytest tests/test_gpu_build.py -m gpu -x -q > gpurun_out/r2z_pytest_build.log 2>&1
tail -3 gpurun_out/r2z_pytest_build.log
for k in 1 2 3; do python variants/r2y_build_step.py; done > gpurun_out/r2z_build_times.log 2>&1
cat gpurun_out/r2z_build_times.log
ECHO_B200_PROFILE=1 python - 2>&1 <<'PY' | grep -v "^\[echo_b200 build\] sweep: 64 " | tee gpurun_out/r2z_build_c5.log
import time
from echorenderer_b200 import build_qbvh_device, scenes, _native
d = scenes.large_scene()
build_qbvh_device(d.triangles[:64], d.spheres[:0])
for _ in range(2):
    t = time.perf_counter(); nodes, depth = build_qbvh_device(d.triangles, d.spheres); print("c5 call %.1f ms" % ((time.perf_counter() - t) * 1e3), len(nodes), depth, _native.last_build(), flush=True)
PY

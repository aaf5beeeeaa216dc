#!/bin/bash
# r2aa: the device light-tree build (csrc/lightbuild.cu): its GPU tests first (with the C4-size timing printed), then the whole GPU suite
set -x
mkdir -p gpurun_out
ECHO_B200_PROFILE=1 timeout 600 python -m pytest tests/test_gpu_build.py -m gpu -x -q -s -k "light" > gpurun_out/r2aa_pytest_light.log 2>&1
tail -25 gpurun_out/r2aa_pytest_light.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2aa_pytest_gpu.log 2>&1
tail -5 gpurun_out/r2aa_pytest_gpu.log

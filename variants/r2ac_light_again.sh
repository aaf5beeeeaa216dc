#!/bin/bash
# r2ac: the device light-tree build again: the relative areas off the chains (CostPass), the whole-sphere test on the cosine
set -x
mkdir -p gpurun_out
ECHO_B200_PROFILE=1 timeout 600 python -m pytest tests/test_gpu_build.py -m gpu -x -q -s -k "light" > gpurun_out/r2ac_pytest_light.log 2>&1
tail -5 gpurun_out/r2ac_pytest_light.log
grep "light tree" gpurun_out/r2ac_pytest_light.log | tail -30

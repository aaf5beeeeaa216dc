#!/bin/bash
# r2ak: the light-build backend with its CUB scratch grown on demand: the whole GPU suite once more (the last GPU call of the round)
set -x
mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -x -q > gpurun_out/r2ak_pytest_gpu.log 2>&1
tail -4 gpurun_out/r2ak_pytest_gpu.log

#!/bin/bash
# r2aj: the final build on two GPUs — the two tests that need a second device in one process, then a short 2-rank default bench line
set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_boundary.py -m gpu -x -q -rs -k "two_devices or multi_device" > gpurun_out/r2aj_pytest_2gpu.log 2>&1; tail -4 gpurun_out/r2aj_pytest_2gpu.log
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 10 --warmup 3 --render-steps 4 --no-cpu-baseline > gpurun_out/r2aj_bench_2gpu.json 2> gpurun_out/r2aj_bench_2gpu.err ) 2>&1 | tail -3
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2aj_bench_2gpu.json'))
print('C2', round(d['value']), 'e2e', round(d['e2e']['value']), 'of ceiling', round(d['e2e']['frac_of_copy_ceiling'], 3), d['e2e']['copy_ceiling']['gbs_all_ranks'])
rec = d['render']['c5']
print('c5', round(rec['value'] / 1e6, 1), 'Msamples/s', round(rec['ms_per_step'], 1), 'ms/step', rec['rank_step_ms'], rec['all_reduce'], 'e2e', round(rec['e2e']['value'] / 1e6, 1))
PY

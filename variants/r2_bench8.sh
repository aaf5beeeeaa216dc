# round 2: the default bench line exactly as the driver launches it at N = 8 (and the reference arm), timed
set -x
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2k_bench_8gpu.json 2> gpurun_out/r2k_bench_8gpu.err ) 2>&1 | tail -3
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2k_bench_8gpu.json'))
print('C2', round(d['value']), 'e2e', round(d['e2e']['value']), 'of ceiling', round(d['e2e']['frac_of_copy_ceiling'], 3), round(d['e2e']['copy_ceiling']['gbs_all_ranks'], 1), 'GB/s all ranks', 'pageable', round(d['e2e_pageable']['value']))
rec = d['render']['c5']
print('c5', round(rec['value'] / 1e6, 1), 'Msamples/s', round(rec['ms_per_step'], 1), 'ms/step', rec['steps'], 'steps', rec['rank_step_ms']['per_rank'], rec['all_reduce']['ms_mean'], 'e2e', round(rec['e2e']['value'] / 1e6, 1), 'frac', round(rec['roofline']['frac'], 3), rec['clocks'])
PY
tail -5 gpurun_out/r2k_bench_8gpu.err

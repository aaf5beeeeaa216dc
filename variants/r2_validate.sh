# round 2: GPU suite + the default bench line as the driver runs it (timed) + on-chip peaks, one GPU
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r2b_pytest.log 2>&1; tail -5 gpurun_out/r2b_pytest.log
python -c "
from echorenderer_b200 import _native
print(_native.measure_peaks(0)); print(_native.measure_peaks(0))" > gpurun_out/r2b_peaks.txt 2>&1
cat gpurun_out/r2b_peaks.txt
( time python bench.py --steps 20 --warmup 5 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err ) 2> gpurun_out/r2b_bench.time
tail -3 gpurun_out/r2b_bench.time; grep "bench " gpurun_out/r2b_bench.err | tail -20
( time python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2b_reference.json 2>/dev/null ) 2>&1 | tail -3
python variants/r2_sweep_c5.py --steps 2 --quick > gpurun_out/r2b_sweep_1gpu.jsonl 2> gpurun_out/r2b_sweep_1gpu.err; cat gpurun_out/r2b_sweep_1gpu.jsonl | cut -c1-300
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2b_bench.json'))
r = d['roofline']
print('C2', round(d['value']), 'Mrays/s frac', round(r['frac'], 3), 'frac_l2', r.get('frac_l2'), 'frac_l1', r.get('frac_l1_sectors'), 'e2e', round(d['e2e']['value']), 'pageable', round(d['e2e_pageable']['value']), 'cpu', d['cpu_baseline']['value'])
for key, rec in d['render'].items():
    if isinstance(rec, dict):
        print(key, round(rec['value'] / 1e6, 1), 'Msamples/s', round(rec['ms_per_step'], 1), 'ms/step B_sample', round(rec['roofline']['algorithmic_bytes_per_sample']), 'frac', round(rec['roofline']['frac'], 3),
              'e2e', round(rec['e2e']['value'] / 1e6, 1), 'cpu', round(rec.get('cpu_baseline', {}).get('value', 0) / 1e6, 2), rec.get('cpu_baseline', {}).get('seconds'), 'allreduce', rec['all_reduce'], rec['rank_step_ms']['mean'])
PY

# round 2: after the persistent scratch + per-handle mutex: boundary / render / instancing tests, the e2e probe, the default bench line
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r2q_pytest.log 2>&1; tail -4 gpurun_out/r2q_pytest.log
python variants/r2_e2e_probe.py 2>&1 | tail -4
( time python bench.py --steps 20 --warmup 5 > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err ) 2> gpurun_out/r2q_bench.time; tail -3 gpurun_out/r2q_bench.time
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2q_bench.json'))
r = d['roofline']
print('C2', round(d['value']), 'Mrays/s frac', round(r['frac'], 3), 'frac_l2_sectors', round(r['frac_l2_sectors'], 3), 'e2e', round(d['e2e']['value']), 'of ceiling', round(d['e2e']['frac_of_copy_ceiling'], 3), 'pageable', round(d['e2e_pageable']['value']), 'cpu', round(d['cpu_baseline']['value'], 1), 'secondary', round(d['secondary']['closest_hit']['mrays_per_s']), round(d['secondary']['occlusion']['mrays_per_s']))
for key, rec in d['render'].items():
    if isinstance(rec, dict):
        print(key, round(rec['value'] / 1e6, 1), 'Msamples/s', round(rec['ms_per_step'], 1), 'ms/step B', round(rec['roofline']['algorithmic_bytes_per_sample']), 'frac', round(rec['roofline']['frac'], 3), 'e2e', round(rec['e2e']['value'] / 1e6, 1), round(rec['e2e']['ms_per_step'], 1), 'cpu', round(rec['cpu_baseline']['value'] / 1e6, 2), round(rec['cpu_baseline']['seconds'], 1))
PY

# round 2: 4 vs 8 pipeline slots of the host-buffer legs, then the GPU suite and the default bench line of the shipped build
set -x
ECHO_B200_LIBRARY=$PWD/variants/lib_slots8.so python bench.py --workload trace --steps 10 --no-cpu-baseline --no-secondary 2>/dev/null > gpurun_out/r2p_trace_slots8.json
python bench.py --workload trace --steps 10 --no-cpu-baseline --no-secondary 2>/dev/null > gpurun_out/r2p_trace_slots4.json
python - <<'PY'
import json
for k in ('slots4', 'slots8'):
    d = json.load(open(f'gpurun_out/r2p_trace_{k}.json'))
    print(k, 'value', round(d['value']), 'e2e pinned', round(d['e2e']['value']), 'of ceiling', round(d['e2e']['frac_of_copy_ceiling'], 3), 'e2e pageable', round(d['e2e_pageable']['value']), round(d['e2e_pageable']['ms_per_step'], 1), 'ms')
PY
python -m pytest tests -m gpu -x -q > gpurun_out/r2p_pytest.log 2>&1; tail -4 gpurun_out/r2p_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
( time python bench.py --steps 20 --warmup 5 > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err ) 2> gpurun_out/r2p_bench.time; tail -3 gpurun_out/r2p_bench.time
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2p_bench.json'))
r = d['roofline']
print('C2', round(d['value']), 'Mrays/s frac', round(r['frac'], 3), 'frac_l2_sectors', round(r['frac_l2_sectors'], 3), 'traffic', r['traffic'], 'e2e', round(d['e2e']['value']), 'pageable', round(d['e2e_pageable']['value']), 'cpu', round(d['cpu_baseline']['value'], 1))
for key, rec in d['render'].items():
    if isinstance(rec, dict):
        print(key, round(rec['value'] / 1e6, 1), 'Msamples/s', round(rec['ms_per_step'], 1), 'ms/step frac', round(rec['roofline']['frac'], 3), 'traffic', rec['roofline']['traffic'], 'e2e', round(rec['e2e']['value'] / 1e6, 1), 'cpu', round(rec['cpu_baseline']['value'] / 1e6, 2), rec['all_reduce'])
PY

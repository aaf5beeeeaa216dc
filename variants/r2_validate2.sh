# round 2: GPU suite on the shipped build (unordered any-hit now default), tail / narrow limit sweep, the default bench line
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r2g_pytest.log 2>&1; tail -4 gpurun_out/r2g_pytest.log
python variants/r2_sweep_c5.py --tail-sweep --scene large --spp 32 --steps 2 > gpurun_out/r2g_tail_c5.jsonl 2>/dev/null
python variants/r2_sweep_c5.py --tail-sweep --scene cornell --width 512 --height 512 --spp 16 --steps 4 > gpurun_out/r2g_tail_c1.jsonl 2>/dev/null
python variants/r2_sweep_c5.py --tail-sweep --scene lights --width 1920 --height 1080 --spp 32 --steps 2 > gpurun_out/r2g_tail_c4.jsonl 2>/dev/null
for f in c5 c1 c4; do echo $f; python -c "
import json
for line in open('gpurun_out/r2g_tail_$f.jsonl'):
    d = json.loads(line); print(' ', d['tag'], d['msamples_per_s'], d['ms_per_step'], d['launches_rank0'])
"; done
( time python bench.py --steps 20 --warmup 5 > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err ) 2> gpurun_out/r2g_bench.time; tail -3 gpurun_out/r2g_bench.time
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2g_bench.json'))
r = d['roofline']
print('C2', round(d['value']), 'Mrays/s frac', round(r['frac'], 3), 'frac_l2', round(r['frac_l2'], 3), 'frac_l2_sectors', round(r['frac_l2_sectors'], 3), 'e2e', round(d['e2e']['value']), 'of ceiling', round(d['e2e']['frac_of_copy_ceiling'], 3), d['e2e']['copy_ceiling'], 'pageable', round(d['e2e_pageable']['value']), 'cpu', d['cpu_baseline']['value'])
print('secondary', d['secondary']['closest_hit']['mrays_per_s'], d['secondary']['occlusion']['mrays_per_s'])
for key, rec in d['render'].items():
    if isinstance(rec, dict):
        print(key, round(rec['value'] / 1e6, 1), 'Msamples/s', round(rec['ms_per_step'], 1), 'ms/step B_sample', round(rec['roofline']['algorithmic_bytes_per_sample']), 'frac', round(rec['roofline']['frac'], 3),
              'e2e', round(rec['e2e']['value'] / 1e6, 1), 'cpu', round(rec.get('cpu_baseline', {}).get('value', 0) / 1e6, 2), rec.get('cpu_baseline', {}).get('seconds'))
PY

# round 2: device tree build (PLOC vs LBVH) tests + C2 on each tree, narrow-limit sweep, then the ncu evidence (variants/r2_ncu.sh)
set -x
python -m pytest tests/test_gpu_build.py -m gpu -x -q 2>&1 | tail -5
ECHO_B200_PROFILE=1 python bench.py --workload trace --tree device --steps 5 --no-cpu-baseline 2> gpurun_out/r2h_trace_ploc.err > gpurun_out/r2h_trace_ploc.json
ECHO_B200_BUILD_ALGORITHM=0 python bench.py --workload trace --tree device --steps 5 --no-cpu-baseline 2>/dev/null > gpurun_out/r2h_trace_lbvh.json
grep "echo_b200 build" gpurun_out/r2h_trace_ploc.err | tail -2
python variants/r2_sweep_c5.py --tail-sweep --scene cornell --width 512 --height 512 --spp 16 --steps 4 > gpurun_out/r2h_tail_c1.jsonl 2>/dev/null
python variants/r2_sweep_c5.py --tail-sweep --scene mixed --width 1920 --height 1080 --spp 64 --steps 2 --bounce-limit 8 > gpurun_out/r2h_tail_c3.jsonl 2>/dev/null
python variants/r2_sweep_c5.py --tail-sweep --scene large --spp 32 --steps 2 > gpurun_out/r2h_tail_c5.jsonl 2>/dev/null
python - <<'PY'
import json
for k in ('ploc', 'lbvh'):
    d = json.load(open(f'gpurun_out/r2h_trace_{k}.json')); r = d['roofline']
    print(k, round(d['value']), 'Mrays/s closest', round(r['mrays_per_s']), 'occl', round(r['occlusion']['mrays_per_s']), 'nodes/query', round(r['visits_per_query']['nodes'], 2), d['config']['tree'])
for f in ('c1', 'c3', 'c5'):
    print(f)
    for line in open(f'gpurun_out/r2h_tail_{f}.jsonl'):
        d = json.loads(line); print('  ', d['tag'], d['msamples_per_s'], d['ms_per_step'], d['launches_rank0'])
PY
bash variants/r2_ncu.sh

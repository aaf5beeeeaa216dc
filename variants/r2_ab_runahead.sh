# round 2: GPU suite on the new run-ahead wavefront loop + self-resetting ray counters, then A/B of the run-ahead depth
# (0 = round 1's lock-step loop) on C1 / C3 / C4 / C5
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; tail -5 gpurun_out/r2a_pytest.log
for ra in 0 1 3 6; do
  export ECHO_B200_RUN_AHEAD=$ra
  python bench.py --workload render --scene cornell --width 512 --height 512 --spp 16 --steps 5 --bounce-limit 128 2>/dev/null > gpurun_out/r2a_c1_ra$ra.json
  python bench.py --workload render --scene mixed --spp 64 --steps 4 --bounce-limit 8 2>/dev/null > gpurun_out/r2a_c3_ra$ra.json
  python bench.py --workload render --scene lights --spp 64 --steps 4 --bounce-limit 128 2>/dev/null > gpurun_out/r2a_c4_ra$ra.json
  python variants/r2_probe_shard.py --world 8 --spp 256 --tag shard8_ra$ra > gpurun_out/r2a_c5_ra$ra.json 2>/dev/null
done
unset ECHO_B200_RUN_AHEAD
python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench_trace.json 2> gpurun_out/r2a_bench_trace.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2a_c*_ra*.json')):
    try:
        d = json.load(open(f))
        print(f, round(d.get('value', 0) / 1e6, 1) if 'value' in d else round(d['msamples_per_s'], 1), 'Msamples/s', round(d['ms_per_step'], 2), 'ms/step', d.get('gpu_launches', d.get('launches_per_step')))
    except Exception as e:
        print(f, 'failed', e)
d = json.load(open('gpurun_out/r2a_bench_trace.json'))
print('trace', d['value'], d['roofline']['mrays_per_s'], d['roofline']['occlusion']['mrays_per_s'], d['e2e']['value'], d['secondary']['closest_hit']['mrays_per_s'])
PY

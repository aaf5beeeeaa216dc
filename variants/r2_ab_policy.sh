# round 2: L2 residency hints (ECHO_RAY_POLICY: evict-first ray staging, ECHO_NODE_POLICY: evict-last node loads) on C2 / C3 / C5,
# and the PLOC search radius on the C2 geometry
set -x
for v in default raypol nodepol bothpol; do
  if [ $v = default ]; then unset ECHO_B200_LIBRARY; else export ECHO_B200_LIBRARY=$PWD/variants/lib_$v.so; fi
  python bench.py --workload trace --steps 10 --no-cpu-baseline 2>/dev/null > gpurun_out/r2i_trace_$v.json
  python bench.py --workload render --scene mixed --spp 64 --steps 4 --bounce-limit 8 --no-cpu-baseline 2>/dev/null > gpurun_out/r2i_c3_$v.json
  python bench.py --workload render --scene large --width 3840 --height 2160 --spp 32 --steps 3 --bounce-limit 128 --no-cpu-baseline 2>/dev/null > gpurun_out/r2i_c5_$v.json
done
unset ECHO_B200_LIBRARY
for r in 4 16 32 64; do
  ECHO_B200_BUILD_PLOC_RADIUS=$r ECHO_B200_PROFILE=1 python bench.py --workload trace --tree device --steps 5 --no-cpu-baseline --no-secondary 2> gpurun_out/r2i_ploc_r$r.err > gpurun_out/r2i_ploc_r$r.json
  grep "echo_b200 build" gpurun_out/r2i_ploc_r$r.err | tail -1
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2i_*.json')):
    try:
        d = json.load(open(f))
        extra = ''
        if d['unit'] == 'Mrays/s':
            r = d['roofline']
            extra = f"closest {r['mrays_per_s']:.0f} occl {r['occlusion']['mrays_per_s']:.0f} nodes/query {r['visits_per_query']['nodes']:.2f}"
            if d.get('secondary'):
                extra += f" secondary {d['secondary']['closest_hit']['mrays_per_s']:.0f} / {d['secondary']['occlusion']['mrays_per_s']:.0f}"
        print(f, round(d['value'] / (1e6 if d['unit'] == 'samples/s' else 1), 1), d['unit'], round(d['ms_per_step'], 2), 'ms/step', extra)
    except Exception as e:
        print(f, 'failed', e)
PY

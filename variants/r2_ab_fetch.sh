# round 2: ECHO_FETCH_VOTE (idle lanes wait to take their next ray until several are waiting): C2 closest / occlusion / secondary, C3
set -x
for v in default fv4w1 fv6w1 fv8w2; do
  if [ $v = default ]; then unset ECHO_B200_LIBRARY; else export ECHO_B200_LIBRARY=$PWD/variants/lib_$v.so; fi
  python bench.py --workload trace --steps 10 --no-cpu-baseline 2>/dev/null > gpurun_out/r2u_trace_$v.json
  python bench.py --workload render --scene mixed --spp 64 --steps 4 --bounce-limit 8 --no-cpu-baseline 2>/dev/null > gpurun_out/r2u_c3_$v.json
done
unset ECHO_B200_LIBRARY
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2u_*.json')):
    d = json.load(open(f))
    extra = ''
    if d['unit'] == 'Mrays/s':
        r = d['roofline']; extra = f"closest {r['mrays_per_s']:.0f} occl {r['occlusion']['mrays_per_s']:.0f} secondary {d['secondary']['closest_hit']['mrays_per_s']:.0f} / {d['secondary']['occlusion']['mrays_per_s']:.0f}"
    print(f, round(d['value'] / (1e6 if d['unit'] == 'samples/s' else 1), 1), d['unit'], round(d['ms_per_step'], 2), extra)
PY

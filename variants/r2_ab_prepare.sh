# round 2: ECHO_PREPARE_RAYS (per-ray constants computed per pool, 32 lanes at a time, parked in the hit slot): parity first, then A/B
set -x
python -m pytest tests/test_gpu_trace.py tests/test_gpu_fuzz.py tests/test_gpu_full_size.py tests/test_gpu_instancing.py tests/test_gpu_render.py tests/test_gpu_boundary.py -m gpu -x -q 2>&1 | tail -4
for v in default noprep; do
  if [ $v = default ]; then unset ECHO_B200_LIBRARY; else export ECHO_B200_LIBRARY=$PWD/variants/lib_$v.so; fi
  python bench.py --workload trace --steps 10 --no-cpu-baseline 2>/dev/null > gpurun_out/r2v_trace_$v.json
  python bench.py --workload render --scene mixed --spp 64 --steps 4 --bounce-limit 8 --no-cpu-baseline 2>/dev/null > gpurun_out/r2v_c3_$v.json
  python bench.py --workload render --scene large --width 3840 --height 2160 --spp 32 --steps 3 --bounce-limit 128 --no-cpu-baseline 2>/dev/null > gpurun_out/r2v_c5_$v.json
  python bench.py --workload trace --instanced --steps 5 --no-cpu-baseline 2>/dev/null > gpurun_out/r2v_inst_$v.json
done
unset ECHO_B200_LIBRARY
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2v_*.json')):
    d = json.load(open(f))
    extra = ''
    if d['unit'] == 'Mrays/s':
        r = d['roofline']; extra = f"closest {r['mrays_per_s']:.0f} occl {r['occlusion']['mrays_per_s']:.0f}"
        if d.get('secondary'): extra += f" secondary {d['secondary']['closest_hit']['mrays_per_s']:.0f} / {d['secondary']['occlusion']['mrays_per_s']:.0f}"
    print(f, round(d['value'] / (1e6 if d['unit'] == 'samples/s' else 1), 1), d['unit'], round(d['ms_per_step'], 2), extra)
PY

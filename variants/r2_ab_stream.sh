# round 2: GPU suite (default build, then the bounds-check build), instanced regression check, A/B of ECHO_STREAM_STATE and ECHO_ANY_UNORDERED
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r2f_pytest.log 2>&1; tail -4 gpurun_out/r2f_pytest.log
ECHO_B200_LIBRARY=$PWD/variants/lib_bounds.so python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_bench.py > gpurun_out/r2f_pytest_bounds.log 2>&1; tail -4 gpurun_out/r2f_pytest_bounds.log
for v in default stream anyunordered; do
  if [ $v = default ]; then unset ECHO_B200_LIBRARY; else export ECHO_B200_LIBRARY=$PWD/variants/lib_$v.so; fi
  python bench.py --workload render --scene mixed --spp 64 --steps 4 --bounce-limit 8 --no-cpu-baseline 2>/dev/null > gpurun_out/r2f_c3_$v.json
  python bench.py --workload render --scene lights --spp 64 --steps 3 --bounce-limit 128 --no-cpu-baseline 2>/dev/null > gpurun_out/r2f_c4_$v.json
  python bench.py --workload render --scene large --width 3840 --height 2160 --spp 32 --steps 3 --bounce-limit 128 --no-cpu-baseline 2>/dev/null > gpurun_out/r2f_c5_$v.json
  python bench.py --workload render --scene instanced --spp 16 --steps 4 --bounce-limit 16 --no-cpu-baseline 2>/dev/null > gpurun_out/r2f_inst_$v.json
  python bench.py --workload trace --steps 10 --no-cpu-baseline 2>/dev/null > gpurun_out/r2f_trace_$v.json
done
unset ECHO_B200_LIBRARY
ECHO_B200_PROFILE=1 python bench.py --workload render --scene large --width 3840 --height 2160 --spp 8 --steps 1 --bounce-limit 128 --no-cpu-baseline 2> gpurun_out/r2f_c5_profile.err > /dev/null
grep "echo_b200 profile" gpurun_out/r2f_c5_profile.err | tail -15
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2f_*.json')):
    try:
        d = json.load(open(f))
        extra = ''
        if d['unit'] == 'Mrays/s':
            extra = f"closest {d['roofline']['mrays_per_s']:.0f} occl {d['roofline']['occlusion']['mrays_per_s']:.0f} secondary {d['secondary']['closest_hit']['mrays_per_s']:.0f} / {d['secondary']['occlusion']['mrays_per_s']:.0f}"
        print(f, round(d['value'] / (1e6 if d['unit'] == 'samples/s' else 1), 1), d['unit'], round(d['ms_per_step'], 2), 'ms/step', d['gpu_launches'], extra)
    except Exception as e:
        print(f, 'failed', e)
PY

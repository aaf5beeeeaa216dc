# round 2: instanced render regressed (295 vs 457 M samples/s in r1h)? A/B against the round-1 tree built beside (variants/_r1_tree)
set -x
for i in 1 2; do
python bench.py --workload render --scene instanced --spp 16 --steps 4 --bounce-limit 16 --no-cpu-baseline --pattern ordered 2>/dev/null > gpurun_out/r2e_inst_new_ordered_$i.json
( cd variants/_r1_tree && python bench.py --workload render --scene instanced --spp 16 --steps 4 --bounce-limit 16 2>/dev/null ) > gpurun_out/r2e_inst_r1_$i.json
done
python bench.py --workload render --scene instanced --spp 16 --steps 4 --bounce-limit 16 --no-cpu-baseline --pattern hilbert 2>/dev/null > gpurun_out/r2e_inst_new_hilbert.json
ECHO_B200_PROFILE=1 python bench.py --workload render --scene instanced --spp 16 --steps 1 --bounce-limit 16 --no-cpu-baseline --pattern ordered 2> gpurun_out/r2e_inst_new_profile.err > /dev/null
( cd variants/_r1_tree && ECHO_B200_PROFILE=1 python bench.py --workload render --scene instanced --spp 16 --steps 1 --bounce-limit 16 2> ../../gpurun_out/r2e_inst_r1_profile.err > /dev/null )
grep "echo_b200 profile" gpurun_out/r2e_inst_new_profile.err | tail -16
grep "echo_b200 profile" gpurun_out/r2e_inst_r1_profile.err | tail -14
( cd variants/_r1_tree && python bench.py --workload render --scene mixed --spp 64 --steps 4 --bounce-limit 8 2>/dev/null ) > gpurun_out/r2e_c3_r1.json
python bench.py --instanced --workload trace --steps 5 --no-cpu-baseline 2>/dev/null > gpurun_out/r2e_trace_inst_new.json
( cd variants/_r1_tree && python bench.py --instanced --steps 5 --no-cpu-baseline 2>/dev/null ) > gpurun_out/r2e_trace_inst_r1.json
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2e_*.json')):
    try:
        d = json.load(open(f)); print(f, round(d['value'] / (1e6 if d['unit'] == 'samples/s' else 1), 1), d['unit'], round(d['ms_per_step'], 2), 'ms/step', d['gpu_launches'])
    except Exception as e:
        print(f, 'failed', e)
PY

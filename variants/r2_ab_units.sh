# round 2: the reordered wavefront (units: shade k | extend + classify k + 1, exact class grids, smooth-dielectric class, CTA-aggregated
# classify) — GPU suite, then C1 / C3 / C4 / C5 and the per-kernel profile of C3
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r2d_pytest.log 2>&1; tail -5 gpurun_out/r2d_pytest.log
ECHO_B200_TAIL_LIMIT=100000000 python -m pytest tests/test_gpu_render.py tests/test_gpu_instancing.py -m gpu -x -q > gpurun_out/r2d_pytest_tail.log 2>&1; tail -3 gpurun_out/r2d_pytest_tail.log
python bench.py --workload render --scene cornell --width 512 --height 512 --spp 16 --steps 5 --bounce-limit 128 --no-cpu-baseline 2>/dev/null > gpurun_out/r2d_c1.json
python bench.py --workload render --scene mixed --spp 64 --steps 4 --bounce-limit 8 --no-cpu-baseline 2>/dev/null > gpurun_out/r2d_c3.json
python bench.py --workload render --scene lights --spp 64 --steps 4 --bounce-limit 128 --no-cpu-baseline 2>/dev/null > gpurun_out/r2d_c4.json
python bench.py --workload render --scene large --width 3840 --height 2160 --spp 32 --steps 3 --bounce-limit 128 --no-cpu-baseline 2>/dev/null > gpurun_out/r2d_c5.json
python bench.py --workload render --scene textured --spp 16 --steps 4 --bounce-limit 16 --no-cpu-baseline 2>/dev/null > gpurun_out/r2d_tex.json
python bench.py --workload render --scene instanced --spp 16 --steps 4 --bounce-limit 16 --no-cpu-baseline 2>/dev/null > gpurun_out/r2d_inst.json
ECHO_B200_PROFILE=1 python bench.py --workload render --scene mixed --spp 16 --steps 1 --bounce-limit 8 --no-cpu-baseline 2> gpurun_out/r2d_c3_profile.err > /dev/null
grep "echo_b200 profile" gpurun_out/r2d_c3_profile.err | tail -16
python - <<'PY'
import json
for k in ['c1', 'c3', 'c4', 'c5', 'tex', 'inst']:
    try:
        d = json.load(open(f'gpurun_out/r2d_{k}.json'))
        print(k, round(d['value'] / 1e6, 1), 'Msamples/s', round(d['ms_per_step'], 2), 'ms/step launches', d['gpu_launches'], 'e2e', round(d['e2e']['value'] / 1e6, 1))
    except Exception as e:
        print(k, 'failed', e)
PY

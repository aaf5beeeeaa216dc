# A/B of ECHO_SHADE_MIN_BLOCKS: resident CTAs per SM asked of the shading kernels (default: no target, 100-110 registers, 4 CTAs)
for v in m6 m7 m8; do
  if [ $v = default ]; then unset ECHO_B200_LIBRARY; else export ECHO_B200_LIBRARY=$PWD/variants/lib_$v.so; fi
  python bench.py --workload render --scene mixed --spp 64 --steps 3 --no-cpu-baseline > gpurun_out/ab15r_$v.json 2>/dev/null
  python bench.py --workload render --scene textured --spp 64 --steps 3 --bounce-limit 16 --no-cpu-baseline > gpurun_out/ab15t_$v.json 2>/dev/null
  python bench.py --workload render --scene lights --spp 64 --steps 2 --bounce-limit 128 --no-cpu-baseline > gpurun_out/ab15l_$v.json 2>/dev/null
  python -c "
import json;r=json.load(open('gpurun_out/ab15r_$v.json'));t=json.load(open('gpurun_out/ab15t_$v.json'));l=json.load(open('gpurun_out/ab15l_$v.json'));print('$v','C3',round(r['value']/1e6,1),'textured',round(t['value']/1e6,1),'C4',round(l['value']/1e6,1))"
  ECHO_B200_PROFILE=1 python bench.py --workload render --scene mixed --spp 16 --steps 1 --no-cpu-baseline 2>&1 >/dev/null | grep -i -E "shade" | tail -5
done

# A/B of ECHO_SHARE_LEVEL (shared, i.e. noinline, helper code in the shading kernels): the default build against a variants/lib_l*.so
for v in level2 level1; do
  if [ $v = level2 ]; then unset ECHO_B200_LIBRARY; else export ECHO_B200_LIBRARY=$PWD/variants/lib_l1.so; fi
  python bench.py --workload render --scene mixed --spp 64 --steps 3 --no-cpu-baseline > gpurun_out/ab13r_$v.json 2>/dev/null
  python bench.py --workload render --scene textured --spp 64 --steps 3 --bounce-limit 16 --no-cpu-baseline > gpurun_out/ab13t_$v.json 2>/dev/null
  python bench.py --workload render --scene cornell --width 512 --height 512 --spp 64 --steps 3 --bounce-limit 128 --no-cpu-baseline > gpurun_out/ab13c_$v.json 2>/dev/null
  python -c "
import json;r=json.load(open('gpurun_out/ab13r_$v.json'));t=json.load(open('gpurun_out/ab13t_$v.json'));c=json.load(open('gpurun_out/ab13c_$v.json'));print('$v','C3',round(r['value']/1e6,1),'textured',round(t['value']/1e6,1),'C1',round(c['value']/1e6,1))"
  ECHO_B200_PROFILE=1 python bench.py --workload render --scene mixed --spp 16 --steps 1 --no-cpu-baseline 2>&1 >/dev/null | grep -i -E "shade|extend|shadow|classify|other" | tail -12
done

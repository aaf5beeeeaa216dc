# A/B of ECHO_SHARED_STACK: the top entries of the traversal stack in shared memory (default build, 8) vs all in local memory (lib_ss0.so)
for v in shared8 local; do
  if [ $v = shared8 ]; then unset ECHO_B200_LIBRARY; else export ECHO_B200_LIBRARY=$PWD/variants/lib_ss0.so; fi
  python bench.py --steps 5 --no-cpu-baseline > gpurun_out/ab14t_$v.json 2>/dev/null
  python bench.py --workload render --scene mixed --spp 64 --steps 3 --no-cpu-baseline > gpurun_out/ab14r_$v.json 2>/dev/null
  python bench.py --workload render --scene lights --spp 64 --steps 2 --bounce-limit 128 --no-cpu-baseline > gpurun_out/ab14l_$v.json 2>/dev/null
  python -c "
import json;d=json.load(open('gpurun_out/ab14t_$v.json'));r=json.load(open('gpurun_out/ab14r_$v.json'));l=json.load(open('gpurun_out/ab14l_$v.json'));ro=d['roofline'];s=d['secondary']
print('$v','C2',round(d['value']),'closest',round(ro['mrays_per_s']),'occl',round(ro['occlusion']['mrays_per_s']),'secondary',round(s['closest_hit']['mrays_per_s']),round(s['occlusion']['mrays_per_s']),'C3',round(r['value']/1e6,1),'C4',round(l['value']/1e6,1))"
done

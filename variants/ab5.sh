# A/B of the prefetch variants on C2 (device-resident closest hit / occlusion) and C3
for v in base pf1 pf2 pf3; do
  if [ $v = base ]; then unset ECHO_B200_LIBRARY; else export ECHO_B200_LIBRARY=$PWD/variants/lib_$v.so; fi
  python bench.py --no-secondary --no-cpu-baseline --steps 5 > gpurun_out/ab5_$v.json 2>/dev/null
  python bench.py --workload render --scene mixed --spp 16 --steps 3 > gpurun_out/ab5r_$v.json 2>/dev/null
  python -c "
import json;d=json.load(open('gpurun_out/ab5_$v.json'));r=json.load(open('gpurun_out/ab5r_$v.json'));print('$v', round(d['value']), 'closest', round(d['roofline']['mrays_per_s']), 'occl', round(d['roofline']['occlusion']['mrays_per_s']), 'C3', round(r['value']/1e6,1))"
done

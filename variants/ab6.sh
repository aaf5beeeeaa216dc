# A/B of the occupancy target of the instanced persistent kernels (closest hit / occlusion Mrays/s on the instanced bench scene)
for v in base inst6 inst7; do
  if [ $v = base ]; then unset ECHO_B200_LIBRARY; else export ECHO_B200_LIBRARY=$PWD/variants/lib_$v.so; fi
  python bench.py --instanced --no-cpu-baseline --steps 4 > gpurun_out/ab6_$v.json 2>/dev/null
  python -c "
import json;d=json.load(open('gpurun_out/ab6_$v.json'));print('$v', round(d['value']), 'closest', round(d['roofline']['mrays_per_s']), 'occl', round(d['roofline']['occlusion']['mrays_per_s']))"
done

for v in default v4 v16 b8 b10; do
  if [ $v = default ]; then unset ECHO_B200_LIBRARY; else export ECHO_B200_LIBRARY=$PWD/variants/lib_$v.so; fi
  python bench.py --no-cpu-baseline --steps 5 > gpurun_out/ab_$v.json 2>/dev/null
  python -c "
import json;d=json.load(open('gpurun_out/ab_$v.json'));r=d['roofline'];print('$v',round(d['value']),round(r['mrays_per_s']),round(r['occlusion']['mrays_per_s']))"
done

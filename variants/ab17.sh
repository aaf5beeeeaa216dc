# A/B of ECHO_MIN_BLOCKS for the C2 batch kernels with the shared-memory stack: 7 CTAs per SM (72 registers) vs 8 (64 registers, 4-52 B of spills)
for v in b7 b8 b7 b8; do
  if [ $v = b7 ]; then unset ECHO_B200_LIBRARY; else export ECHO_B200_LIBRARY=$PWD/variants/lib_$v.so; fi
  python bench.py --steps 5 --no-cpu-baseline > gpurun_out/ab17_$v.json 2>/dev/null
  python -c "
import json;d=json.load(open('gpurun_out/ab17_$v.json'));ro=d['roofline'];s=d['secondary'];print('$v','C2',round(d['value']),'closest',round(ro['mrays_per_s']),'occl',round(ro['occlusion']['mrays_per_s']),'secondary',round(s['closest_hit']['mrays_per_s']),round(s['occlusion']['mrays_per_s']))"
done

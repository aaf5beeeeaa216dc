for v in default w1 w2 v1 v4 v16; do
  if [ $v = default ]; then unset ECHO_B200_LIBRARY; else export ECHO_B200_LIBRARY=$PWD/variants/lib_$v.so; fi
  python bench.py --no-cpu-baseline --steps 5 > gpurun_out/ab_$v.json 2>/dev/null
  python bench.py --workload render --scene mixed --spp 16 --steps 4 > gpurun_out/abr_$v.json 2>/dev/null
  python bench.py --workload render --scene lights --spp 16 --steps 3 --bounce-limit 128 > gpurun_out/abl_$v.json 2>/dev/null
  python -c "
import json;d=json.load(open('gpurun_out/ab_$v.json'));r=d['roofline'];e=json.load(open('gpurun_out/abr_$v.json'));l=json.load(open('gpurun_out/abl_$v.json'));print('$v',round(r['mrays_per_s']),round(r['occlusion']['mrays_per_s']),'C3',round(e['value']/1e6,1),'C4',round(l['value']/1e6,1))"
done

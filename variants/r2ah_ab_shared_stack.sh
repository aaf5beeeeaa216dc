#!/bin/bash
# r2ah: entries of the traversal stack kept in shared memory (ECHO_SHARED_STACK): 8 (shipped) vs 4 / 12 / 16, C2 batch + secondary batch
mkdir -p gpurun_out
for v in base ss4 ss12 ss16 base; do
  if [ $v = base ]; then unset ECHO_B200_LIBRARY; else export ECHO_B200_LIBRARY=$PWD/variants/lib_$v.so; fi
  python bench.py --workload trace --steps 10 --no-cpu-baseline --no-tree-build > gpurun_out/r2ah_$v.json 2>/dev/null
  python -c "
import json;d=json.load(open('gpurun_out/r2ah_$v.json'));ro=d['roofline'];s=d['secondary'];print('$v','C2',round(d['value']),'closest',round(ro['mrays_per_s']),'occl',round(ro['occlusion']['mrays_per_s']),'secondary',round(s['closest_hit']['mrays_per_s']),round(s['occlusion']['mrays_per_s']))"
done

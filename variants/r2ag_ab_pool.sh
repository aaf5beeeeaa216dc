#!/bin/bash
# r2ag: the ray pool a warp reserves per global atomic, 256 (shipped in r2ae) vs 128 / 64 / 32 in every persistent traversal kernel
# (batch, instanced batch, extend, shadow): C2 + the secondary batch, C3 and C4
mkdir -p gpurun_out
for v in base p128 p64 p32 base; do
  if [ $v = base ]; then unset ECHO_B200_LIBRARY; else export ECHO_B200_LIBRARY=$PWD/variants/lib_$v.so; fi
  python bench.py --workload trace --steps 10 --no-cpu-baseline --no-tree-build > gpurun_out/r2ag_$v.json 2>/dev/null
  python bench.py --workload render --scene mixed --spp 64 --steps 4 --bounce-limit 8 --no-cpu-baseline --no-tree-build 2>/dev/null > gpurun_out/r2ag_c3_$v.json
  python bench.py --workload render --scene lights --spp 64 --steps 3 --bounce-limit 128 --no-cpu-baseline --no-tree-build 2>/dev/null > gpurun_out/r2ag_c4_$v.json
  python -c "
import json;d=json.load(open('gpurun_out/r2ag_$v.json'));ro=d['roofline'];s=d['secondary'];print('$v','C2',round(d['value']),'closest',round(ro['mrays_per_s']),'occl',round(ro['occlusion']['mrays_per_s']),'secondary',round(s['closest_hit']['mrays_per_s']),round(s['occlusion']['mrays_per_s']), 'C3', round(json.load(open('gpurun_out/r2ag_c3_$v.json'))['value']/1e6,1), 'C4', round(json.load(open('gpurun_out/r2ag_c4_$v.json'))['value']/1e6,1))"
done

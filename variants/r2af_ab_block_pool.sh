#!/bin/bash
# r2af: A/B of the persistent traversal kernels' CTA size (64 / 128 / 256 threads at the same 896 / 768 resident threads per SM), of the ray pool
# a warp reserves per global atomic (128 / 256 / 512) and of 6 CTAs per SM (80 registers) — the C2 batch and the secondary-ray batch
mkdir -p gpurun_out
for v in base b64 b256 p128 p512 m6 base; do
  if [ $v = base ]; then unset ECHO_B200_LIBRARY; else export ECHO_B200_LIBRARY=$PWD/variants/lib_$v.so; fi
  python bench.py --workload trace --steps 10 --no-cpu-baseline --no-tree-build > gpurun_out/r2af_$v.json 2>/dev/null
  python -c "
import json;d=json.load(open('gpurun_out/r2af_$v.json'));ro=d['roofline'];s=d['secondary'];print('$v','C2',round(d['value']),'closest',round(ro['mrays_per_s']),'occl',round(ro['occlusion']['mrays_per_s']),'secondary',round(s['closest_hit']['mrays_per_s']),round(s['occlusion']['mrays_per_s']))"
done

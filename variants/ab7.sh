# A/B of CUDA graphs for the narrow wavefront iterations (ECHO_B200_GRAPHS=0 launches the nine kernels one by one)
for g in 0 1; do
  export ECHO_B200_GRAPHS=$g
  python bench.py --workload render --scene cornell --width 512 --height 512 --spp 16 --steps 4 --bounce-limit 128 > gpurun_out/ab7c_$g.json 2>/dev/null
  python bench.py --workload render --scene mixed --spp 16 --steps 4 > gpurun_out/ab7r_$g.json 2>/dev/null
  python bench.py --workload render --scene lights --spp 16 --steps 3 --bounce-limit 128 > gpurun_out/ab7l_$g.json 2>/dev/null
  python bench.py --workload render --scene large --width 3840 --height 2160 --spp 16 --steps 2 --bounce-limit 128 > gpurun_out/ab7g_$g.json 2>/dev/null
  python -c "
import json;c=json.load(open('gpurun_out/ab7c_$g.json'));e=json.load(open('gpurun_out/ab7r_$g.json'));l=json.load(open('gpurun_out/ab7l_$g.json'));g=json.load(open('gpurun_out/ab7g_$g.json'));print('graphs $g','C1',round(c['value']/1e6,1),'C3',round(e['value']/1e6,1),'C4',round(l['value']/1e6,1),'C5',round(g['value']/1e6,1))"
done

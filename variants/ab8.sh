# A/B of the tail kernel threshold (live paths below which one kernel finishes them)
for t in 0 4096 16384 65536 262144; do
  export ECHO_B200_TAIL_LIMIT=$t
  python bench.py --workload render --scene cornell --width 512 --height 512 --spp 16 --steps 4 --bounce-limit 128 > gpurun_out/ab8c_$t.json 2>/dev/null
  python bench.py --workload render --scene lights --spp 16 --steps 3 --bounce-limit 128 > gpurun_out/ab8l_$t.json 2>/dev/null
  python bench.py --workload render --scene large --width 3840 --height 2160 --spp 16 --steps 2 --bounce-limit 128 > gpurun_out/ab8g_$t.json 2>/dev/null
  python bench.py --workload render --scene mixed --spp 16 --steps 3 > gpurun_out/ab8r_$t.json 2>/dev/null
  python -c "
import json;c=json.load(open('gpurun_out/ab8c_$t.json'));l=json.load(open('gpurun_out/ab8l_$t.json'));g=json.load(open('gpurun_out/ab8g_$t.json'));r=json.load(open('gpurun_out/ab8r_$t.json'));print('tail $t','C1',round(c['value']/1e6,1),'C3',round(r['value']/1e6,1),'C4',round(l['value']/1e6,1),'C5',round(g['value']/1e6,1), 'launches C1', c['gpu_launches'])"
done

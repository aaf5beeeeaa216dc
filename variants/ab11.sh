# A/B of the number of pipelines with 16 Mi-path batches, 64 spp per step
for w in 8 12 16; do
  export ECHO_B200_RENDER_WORKERS=$w
  python bench.py --workload render --scene mixed --spp 64 --steps 2 --no-cpu-baseline > gpurun_out/ab11r_$w.json 2>/dev/null
  python bench.py --workload render --scene large --width 3840 --height 2160 --spp 32 --steps 2 --bounce-limit 128 --no-cpu-baseline > gpurun_out/ab11g_$w.json 2>/dev/null
  python -c "
import json;g=json.load(open('gpurun_out/ab11g_$w.json'));r=json.load(open('gpurun_out/ab11r_$w.json'));print('workers $w','C3',round(r['value']/1e6,1),'C5',round(g['value']/1e6,1), 'launches', r['gpu_launches'], g['gpu_launches'])"
done

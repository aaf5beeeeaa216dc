# A/B of the wavefront batch size (paths per batch and pipeline) at a realistic 64 spp per step
for b in 4194304 8388608 16777216; do
  export ECHO_B200_BATCH_PATHS=$b
  python bench.py --workload render --scene mixed --spp 64 --steps 2 --no-cpu-baseline > gpurun_out/ab10r_$b.json 2>/dev/null
  python bench.py --workload render --scene lights --spp 64 --steps 2 --bounce-limit 128 --no-cpu-baseline > gpurun_out/ab10l_$b.json 2>/dev/null
  python bench.py --workload render --scene large --width 3840 --height 2160 --spp 16 --steps 2 --bounce-limit 128 --no-cpu-baseline > gpurun_out/ab10g_$b.json 2>/dev/null
  python -c "
import json;l=json.load(open('gpurun_out/ab10l_$b.json'));g=json.load(open('gpurun_out/ab10g_$b.json'));r=json.load(open('gpurun_out/ab10r_$b.json'));print('batch $b','C3',round(r['value']/1e6,1),'C4',round(l['value']/1e6,1),'C5',round(g['value']/1e6,1), 'launches', r['gpu_launches'], l['gpu_launches'], g['gpu_launches'])"
done

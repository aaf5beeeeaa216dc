set -x
python bench.py --workload render --scene cornell --width 512 --height 512 --spp 16 --steps 3 --bounce-limit 128 2>/dev/null > gpurun_out/cfg_c1.json
python bench.py --workload render --scene mixed --spp 64 --steps 4 --bounce-limit 8 2>/dev/null > gpurun_out/cfg_c3.json
python bench.py --workload render --scene lights --spp 64 --steps 4 --bounce-limit 128 2>/dev/null > gpurun_out/cfg_c4.json
python bench.py --workload render --scene large --width 3840 --height 2160 --spp 16 --steps 3 --bounce-limit 128 2>gpurun_out/cfg_c5.err > gpurun_out/cfg_c5.json
for c in c1 c3 c4 c5; do python -c "
import json;d=json.load(open('gpurun_out/cfg_$c.json'));print('$c',round(d['value']/1e6,1),'Msamples/s',round(d['ms_per_step'],1),'ms/step e2e',round(d['e2e']['value']/1e6,1),d['stats_last_step']['Bounce/Created'],d['gpu_launches'])"; done
tail -3 gpurun_out/cfg_c5.err

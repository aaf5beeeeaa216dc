# A/B of the tile sequence handed to the renderer (EvaluationProfile.Pattern): the reference's default Hilbert curve vs row-major
for p in hilbert ordered; do
  python bench.py --workload render --scene mixed --spp 16 --steps 3 --pattern $p --no-cpu-baseline > gpurun_out/ab9r_$p.json 2>/dev/null
  python bench.py --workload render --scene lights --spp 16 --steps 3 --bounce-limit 128 --pattern $p --no-cpu-baseline > gpurun_out/ab9l_$p.json 2>/dev/null
  python bench.py --workload render --scene large --width 3840 --height 2160 --spp 16 --steps 2 --bounce-limit 128 --pattern $p --no-cpu-baseline > gpurun_out/ab9g_$p.json 2>/dev/null
  python -c "
import json;l=json.load(open('gpurun_out/ab9l_$p.json'));g=json.load(open('gpurun_out/ab9g_$p.json'));r=json.load(open('gpurun_out/ab9r_$p.json'));print('pattern $p','C3',round(r['value']/1e6,1),'C4',round(l['value']/1e6,1),'C5',round(g['value']/1e6,1))"
done

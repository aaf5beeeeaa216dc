for w in 0 16384 65536 262144 1048576; do
  export ECHO_B200_NARROW_LIMIT=$w
  python bench.py --workload render --scene cornell --width 512 --height 512 --spp 16 --steps 4 --bounce-limit 128 > gpurun_out/abc_$w.json 2>/dev/null
  python bench.py --workload render --scene mixed --spp 16 --steps 4 > gpurun_out/abr_$w.json 2>/dev/null
  python bench.py --workload render --scene lights --spp 16 --steps 3 --bounce-limit 128 > gpurun_out/abl_$w.json 2>/dev/null
  python bench.py --workload render --scene large --width 3840 --height 2160 --spp 16 --steps 2 --bounce-limit 128 > gpurun_out/abg_$w.json 2>/dev/null
  python -c "
import json;c=json.load(open('gpurun_out/abc_$w.json'));e=json.load(open('gpurun_out/abr_$w.json'));l=json.load(open('gpurun_out/abl_$w.json'));g=json.load(open('gpurun_out/abg_$w.json'));print('narrow $w','C1',round(c['value']/1e6,1),'C3',round(e['value']/1e6,1),'C4',round(l['value']/1e6,1),'C5',round(g['value']/1e6,1))"
done

# A/B of the host-buffer pipeline's chunk size (rays per H2D -> kernel -> D2H chunk), e2e leg of the C2 bench
for c in 262144 524288 1048576 2097152 4194304; do
  ECHO_B200_CHUNK_RAYS=$c python bench.py --steps 3 --no-cpu-baseline --no-secondary > gpurun_out/ab16_$c.json 2>/dev/null
  python -c "
import json;d=json.load(open('gpurun_out/ab16_$c.json'));print('chunk $c', 'e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],2), 'device', round(d['value']))"
done

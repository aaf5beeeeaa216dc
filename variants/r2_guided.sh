# round 2: guided batch sizing (batches shrink towards the end of a call) on one GPU: rank 0's eighth of C5 (what a rank of the 8-GPU job renders),
# the whole C5 frame at 32 spp, C3, C1; then the suite's render tests
set -x
for g in 1 0; do
  export ECHO_B200_GUIDED_BATCHES=$g
  python variants/r2_probe_shard.py --world 8 --spp 256 --pattern hilbert --by position --steps 4 --tag shard8_guided$g > gpurun_out/r2s_shard8_g$g.json 2>/dev/null
  python variants/r2_probe_shard.py --world 1 --spp 32 --pattern hilbert --steps 3 --tag full32_guided$g > gpurun_out/r2s_full32_g$g.json 2>/dev/null
  python variants/r2_probe_shard.py --scene mixed --width 1920 --height 1080 --world 1 --spp 64 --bounce-limit 8 --pattern hilbert --steps 4 --tag c3_guided$g > gpurun_out/r2s_c3_g$g.json 2>/dev/null
  python variants/r2_probe_shard.py --scene cornell --width 512 --height 512 --world 1 --spp 16 --pattern hilbert --steps 6 --tag c1_guided$g > gpurun_out/r2s_c1_g$g.json 2>/dev/null
  python variants/r2_probe_shard.py --scene lights --width 1920 --height 1080 --world 1 --spp 64 --pattern hilbert --steps 3 --tag c4_guided$g > gpurun_out/r2s_c4_g$g.json 2>/dev/null
done
unset ECHO_B200_GUIDED_BATCHES
python -m pytest tests/test_gpu_render.py tests/test_gpu_full_size.py tests/test_gpu_boundary.py -m gpu -x -q 2>&1 | tail -3
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2s_*.json')):
    d = json.load(open(f)); print(d['tag'], round(d['ms_per_step'], 1), 'ms', round(d['msamples_per_s'], 1), 'M/s', d['launches_per_step'])
PY

# round 2: wait mode x run-ahead x pipelines at 8 GPUs on the C5 step (one torchrun session), after the wavefront was reordered into units
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 variants/r2_sweep_c5.py --wait-sweep --steps 3 > gpurun_out/r2l_sweep_wait_8gpu.jsonl 2> gpurun_out/r2l_sweep_wait_8gpu.err
python - <<'PY'
import json
for line in open('gpurun_out/r2l_sweep_wait_8gpu.jsonl'):
    d = json.loads(line); print(d['tag'], '|', d['ms_per_step'], 'ms |', d['msamples_per_s'], 'M/s | ranks', min(d['rank_render_ms_mean']), '-', max(d['rank_render_ms_mean']), '| allreduce', d['all_reduce_ms'])
PY
tail -3 gpurun_out/r2l_sweep_wait_8gpu.err

# Host-core pressure, emulating the 8-GPU box (32 cores / 8 ranks = 4 cores per rank) on two GPUs with taskset: pipelines per rank and
# blocking vs spinning waits. C4 (bounce limit 128: many narrow iterations), 128 spp per step.
run() { # name cores workers blocking
  ECHO_B200_RENDER_WORKERS=$3 ECHO_B200_BLOCKING_SYNC=$4 taskset -c $2 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus 2 --workload render --scene lights --spp 128 --steps 2 --bounce-limit 128 --no-cpu-baseline > gpurun_out/ab12_$1.json 2>/dev/null
  python -c "
import json;d=json.load(open('gpurun_out/ab12_$1.json'));print('$1', round(d['value']/1e6,1), round(d['ms_per_step'],1), round(d['e2e']['value']/1e6,1))"
}
nproc
run free_w8_spin 0-63 8 0
run c8_w8_block 0-7 8 1
run c8_w8_spin 0-7 8 0
run c8_w4_spin 0-7 4 0
run c8_w4_block 0-7 4 1
run c8_w2_spin 0-7 2 0

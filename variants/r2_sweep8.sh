# round 2: the C5 step at 8 GPUs under a sweep of host-side switches (one torchrun session), see r2_sweep_c5.py
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 variants/r2_sweep_c5.py --steps 3 > gpurun_out/r2c_sweep_8gpu.jsonl 2> gpurun_out/r2c_sweep_8gpu.err
cat gpurun_out/r2c_sweep_8gpu.jsonl | cut -c1-420
tail -5 gpurun_out/r2c_sweep_8gpu.err

# round 2, first GPU call: where do the ~90 ms per rank of the 8-GPU C5 step come from? (see r2_probe_shard.py)
set -x
python variants/r2_probe_shard.py --world 1 --spp 32 --tag full_frame_32spp > gpurun_out/r2p_full32.json 2> gpurun_out/r2p_full32.err
python variants/r2_probe_shard.py --world 8 --spp 256 --tag shard8_256spp > gpurun_out/r2p_shard8.json 2> gpurun_out/r2p_shard8.err
ECHO_B200_BATCH_PATHS=8388608 python variants/r2_probe_shard.py --world 8 --spp 256 --tag shard8_256spp_8Mi > gpurun_out/r2p_shard8_8Mi.json 2>/dev/null
ECHO_B200_BATCH_PATHS=4194304 python variants/r2_probe_shard.py --world 8 --spp 256 --tag shard8_256spp_4Mi > gpurun_out/r2p_shard8_4Mi.json 2>/dev/null
ECHO_B200_BATCH_PATHS=2097152 python variants/r2_probe_shard.py --world 8 --spp 256 --tag shard8_256spp_2Mi > gpurun_out/r2p_shard8_2Mi.json 2>/dev/null
ECHO_B200_RENDER_WORKERS=16 ECHO_B200_BATCH_PATHS=4194304 python variants/r2_probe_shard.py --world 8 --spp 256 --tag shard8_256spp_4Mi_w16 > gpurun_out/r2p_shard8_4Mi_w16.json 2>/dev/null
python variants/r2_probe_shard.py --world 8 --spp 256 --pattern hilbert --tag shard8_hilbert > gpurun_out/r2p_shard8_hilbert.json 2>/dev/null
cat gpurun_out/r2p_*.json

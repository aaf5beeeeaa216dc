# round 2: the ncu evidence for the shipped build (one gpurun call; every ncu run follows a plain run of the same command)
#   1. launch list of the C2 command (shares of the step) and `--set full` of its two kernels (DRAM traffic per launch -> profiles/ncu_traffic.json)
#   2. launch list + DRAM traffic of ONE C3 step (16 spp, one pipeline) and `--set full` of the kernels of its second bounce
set -x
TRACE="python bench.py --workload trace --steps 2 --warmup 3 --no-cpu-baseline --no-secondary"
$TRACE > gpurun_out/r2n_trace_plain.json 2> gpurun_out/r2n_trace_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r2n_launches_trace.csv $TRACE > gpurun_out/r2n_trace_ncu1.log 2>&1
$TRACE > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:persistent_batch_kernel -s 2 -c 2 -f -o gpurun_out/r2n_prof_trace $TRACE > gpurun_out/r2n_trace_ncu2.log 2>&1
STEP="python variants/r2_ncu_step.py --scene mixed --spp 16"
$STEP > gpurun_out/r2n_c3_plain.json 2> gpurun_out/r2n_c3_plain.err &&
ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2n_launches_c3.csv $STEP > gpurun_out/r2n_c3_ncu1.log 2>&1
$STEP > /dev/null 2>&1 &&
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"extend_kernel|classify_kernel|shade_kernel|shadow_kernel" -s 9 -c 9 -f -o gpurun_out/r2n_prof_c3 $STEP > gpurun_out/r2n_c3_ncu2.log 2>&1
ls -la gpurun_out/r2n_*; tail -3 gpurun_out/r2n_c3_ncu2.log

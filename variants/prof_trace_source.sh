# source-level (SASS) ncu of the C2 closest-hit kernel
CMD="python bench.py --steps 1 --no-cpu-baseline --no-secondary"
$CMD > gpurun_out/prof_trace_plain.json 2> gpurun_out/prof_trace_plain.err && \
ncu --set full --clock-control none --import-source on -k regex:'persistent_batch_kernel' -s 4 -c 1 \
  -o /tmp/prof_trace -f $CMD > gpurun_out/prof_trace_ncu.log 2>&1
echo rc=$?
ncu -i /tmp/prof_trace.ncu-rep --page source --csv > gpurun_out/prof_trace_source.csv
ncu -i /tmp/prof_trace.ncu-rep --page details | grep -E "persistent_batch_kernel|Duration|Registers Per|Theoretical Occ|Achieved Occ|No Eligible|Issued Warp|Executed Ipc A" | head -12

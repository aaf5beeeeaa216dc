# ncu --set full of the wavefront's kernels on C3 (half-resolution frame, one pipeline so that launches arrive in order): the second
# bounce of the first batch (8 launches: extend, classify, shade x 5, shadow). Only the CSV pages travel back (the report is > 64 MiB).
export ECHO_B200_RENDER_WORKERS=1
CMD="python bench.py --workload render --scene mixed --width 960 --height 540 --spp 16 --steps 1 --no-cpu-baseline"
$CMD > gpurun_out/prof_wavefront_plain.json 2> gpurun_out/prof_wavefront_plain.err && \
ncu --set full --clock-control none -k regex:'extend_kernel|shadow_kernel|shade_kernel|classify_kernel' -s 8 -c 8 \
  -o /tmp/prof_wavefront -f $CMD > gpurun_out/prof_wavefront_ncu.log 2>&1
echo rc=$?
ncu -i /tmp/prof_wavefront.ncu-rep --page raw --csv > gpurun_out/prof_wavefront_raw.csv
ncu -i /tmp/prof_wavefront.ncu-rep --page details > gpurun_out/prof_wavefront_details.txt
ls -la gpurun_out/

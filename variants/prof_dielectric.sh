# source-level ncu of shade_kernel<dielectric> (second bounce of the first batch), C3 at 960 x 540, one pipeline
export ECHO_B200_RENDER_WORKERS=1
CMD="python bench.py --workload render --scene mixed --width 960 --height 540 --spp 16 --steps 1 --no-cpu-baseline"
$CMD > gpurun_out/prof_dielectric_plain.json 2> gpurun_out/prof_dielectric_plain.err && \
ncu --set full --clock-control none --import-source on -k regex:'shade_kernel' -s 7 -c 1 \
  -o /tmp/prof_dielectric -f $CMD > gpurun_out/prof_dielectric_ncu.log 2>&1
echo rc=$?
ncu -i /tmp/prof_dielectric.ncu-rep --page source --csv > gpurun_out/prof_dielectric_source.csv
ncu -i /tmp/prof_dielectric.ncu-rep --page details | head -60
ls -la /tmp/prof_dielectric.ncu-rep gpurun_out/

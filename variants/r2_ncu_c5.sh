# round 2: is C5's traversal (630 MB of nodes, 5x the L2) DRAM-bound? ncu --set full of extend / shadow of its second and third bounce, one pipeline
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r2t_pytest.log 2>&1; tail -3 gpurun_out/r2t_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
STEP="python variants/r2_ncu_step.py --scene large --width 3840 --height 2160 --spp 4 --bounce-limit 128"
$STEP > gpurun_out/r2t_c5_plain.json 2> gpurun_out/r2t_c5_plain.err &&
ncu --profile-from-start off --set full --clock-control none -k regex:"extend_kernel|shadow_kernel" -s 1 -c 4 -f -o gpurun_out/r2t_prof_c5 $STEP > gpurun_out/r2t_c5_ncu.log 2>&1
tail -2 gpurun_out/r2t_c5_ncu.log; cat gpurun_out/r2t_c5_plain.json

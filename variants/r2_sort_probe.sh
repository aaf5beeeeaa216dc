# round 2: upper bound of what sorting the ray queue would buy on the secondary-ray batch, and the ncu capture of that batch's kernels
set -x
python variants/r2_sort_probe.py > gpurun_out/r2m_sort_probe.jsonl 2> gpurun_out/r2m_sort_probe.err && cat gpurun_out/r2m_sort_probe.jsonl &&
ncu --set full --clock-control none --import-source on -k regex:persistent_batch_kernel -s 24 -c 2 -f -o gpurun_out/r2m_prof_secondary python variants/r2_sort_probe.py --steps 2 > gpurun_out/r2m_ncu.log 2>&1
tail -3 gpurun_out/r2m_ncu.log

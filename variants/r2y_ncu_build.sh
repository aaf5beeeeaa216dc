#!/bin/bash
# r2y: ncu launch list of one device SweepBuilder call (C2 geometry): which passes the 15 ms go to
set -x
STEP="python variants/r2y_build_step.py"
$STEP > gpurun_out/r2y_build_plain.log 2> gpurun_out/r2y_build_plain.err &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2y_launches_build.csv $STEP > gpurun_out/r2y_build_ncu.log 2>&1
cat gpurun_out/r2y_build_plain.log; tail -2 gpurun_out/r2y_build_ncu.log; wc -l gpurun_out/r2y_launches_build.csv

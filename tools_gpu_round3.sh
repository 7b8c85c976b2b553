#!/bin/bash
mkdir -p gpurun_out
python tools/run_step.py 32 3 > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/launches_r1.csv python tools/run_step.py 32 3 > gpurun_out/ncu.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/plain.log gpurun_out/ncu.log; wc -l gpurun_out/launches_r1.csv

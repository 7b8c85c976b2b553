#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -x > gpurun_out/tests4.log 2>&1
echo "tests exit $?" | tee gpurun_out/summary4.txt
tail -12 gpurun_out/tests4.log
timeout 600 python tools/profile_plan.py 32 > gpurun_out/plan_B32.csv 2> gpurun_out/plan_B32.err; echo "profile exit $?" | tee -a gpurun_out/summary4.txt; tail -3 gpurun_out/plan_B32.err
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" | tee -a gpurun_out/summary4.txt
cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err

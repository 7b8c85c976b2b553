#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sampler_gpu.py -q -m gpu > gpurun_out/sampler.log 2>&1
echo "sampler exit $?" | tee gpurun_out/summary2.txt
tail -30 gpurun_out/sampler.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" | tee -a gpurun_out/summary2.txt; tail -3 gpurun_out/smoke.log
timeout 600 python tools/profile_plan.py 32 > gpurun_out/plan_B32.csv 2> gpurun_out/plan_B32.err; echo "profile exit $?" | tee -a gpurun_out/summary2.txt; tail -3 gpurun_out/plan_B32.err
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" | tee -a gpurun_out/summary2.txt
cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err

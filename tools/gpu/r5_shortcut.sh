#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_sampler_gpu.py -q -m gpu -x -k shortcut > gpurun_out/tests5.log 2>&1
echo "tests exit $?"; tail -3 gpurun_out/tests5.log
timeout 600 python tools/profile_plan.py 32 > gpurun_out/plan_B32.csv 2> gpurun_out/plan_B32.err; echo "profile exit $?"; tail -3 gpurun_out/plan_B32.err

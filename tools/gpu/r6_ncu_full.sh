#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "gn_silu or conv_k3 or time_mlp" > gpurun_out/tests6.log 2>&1
echo "tests exit $?"; tail -3 gpurun_out/tests6.log
python tools/run_step.py 32 2 > gpurun_out/plain.log 2>&1 || exit 1
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:conv_gemm -s 19 -c 1 -f -o gpurun_out/conv_k768 python tools/run_step.py 32 2 > gpurun_out/ncu1.log 2>&1
$NCU -k regex:conv_gemm -s 35 -c 1 -f -o gpurun_out/conv_k3584 python tools/run_step.py 32 2 > gpurun_out/ncu2.log 2>&1
$NCU -k regex:gn_silu -s 0 -c 1 -f -o gpurun_out/gn_l0 python tools/run_step.py 32 2 > gpurun_out/ncu3.log 2>&1
$NCU -k regex:cross_attn -s 0 -c 1 -f -o gpurun_out/attn_l0 python tools/run_step.py 32 2 > gpurun_out/ncu4.log 2>&1
ls -la gpurun_out/*.ncu-rep
timeout 600 python tools/profile_plan.py 32 > gpurun_out/plan_B32.csv 2> gpurun_out/plan_B32.err; tail -2 gpurun_out/plan_B32.err

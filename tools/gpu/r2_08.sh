#!/bin/bash
O=gpurun_out/r2_08; mkdir -p $O; rm -f gpurun_out/test_metrics.jsonl
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "conv or gn or bias" > $O/conv.log 2>&1; rc=$?
echo "conv tests exit $rc" | tee $O/summary.txt; tail -3 $O/conv.log
if [ $rc -ne 0 ]; then exit 0; fi
timeout 600 python tools/profile_plan.py 32 > $O/plan_B32.csv 2> $O/plan_B32.err; echo "plan exit $?" | tee -a $O/summary.txt; tail -3 $O/plan_B32.err
timeout 900 python -m pytest tests -q -m gpu -x --deselect tests/test_fullsize_gpu.py > $O/tests.log 2>&1; echo "tests exit $?" | tee -a $O/summary.txt; tail -5 $O/tests.log
timeout 900 python bench.py --steps 50 --warmup 5 --no-cpu > $O/bench.json 2> $O/bench.err; echo "bench exit $?" | tee -a $O/summary.txt; cut -c1-300 $O/bench.json; tail -3 $O/bench.err
python tools/bench_conv.py ncu "l3 conv1 1024 k3" > $O/plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_gemm -s 12 -c 1 -f -o $O/conv_xf_l3 python tools/bench_conv.py ncu "l3 conv1 1024 k3" > $O/ncu.log 2>&1; echo "ncu exit $?"; tail -2 $O/ncu.log

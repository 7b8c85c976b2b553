#!/bin/bash
O=gpurun_out/r15; mkdir -p $O
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "conv" > $O/conv.log 2>&1; rc=$?; echo "conv exit $rc" | tee $O/summary.txt; tail -25 $O/conv.log
if [ $rc -ne 0 ]; then exit 0; fi
timeout 900 python -m pytest tests -q -m gpu -x > $O/tests.log 2>&1; echo "tests exit $?" | tee -a $O/summary.txt; tail -4 $O/tests.log
timeout 600 python tools/bench_conv.py > $O/conv_bench.txt 2>&1; cat $O/conv_bench.txt
timeout 600 python tools/profile_plan.py 32 > $O/plan_auto.csv 2> $O/plan_auto.err; tail -2 $O/plan_auto.err
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu > $O/bench.json 2> $O/bench.err; echo "bench exit $?" | tee -a $O/summary.txt; cut -c1-250 $O/bench.json

#!/bin/bash
O=gpurun_out/r27; mkdir -p $O
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "fused_groupnorm or conv" > $O/conv.log 2>&1; rc=$?; echo "conv exit $rc" | tee $O/summary.txt; tail -5 $O/conv.log
if [ $rc -ne 0 ]; then exit 0; fi
timeout 900 python -m pytest tests -q -m gpu -x > $O/tests.log 2>&1; echo "tests exit $?" | tee -a $O/summary.txt; tail -3 $O/tests.log
timeout 900 python bench.py --steps 50 --warmup 5 --no-cpu > $O/bench.json 2> $O/bench.err; echo "bench exit $?" | tee -a $O/summary.txt; cut -c1-240 $O/bench.json; tail -3 $O/bench.err
timeout 600 python tools/profile_plan.py 32 > $O/plan.csv 2> $O/plan.err; tail -2 $O/plan.err

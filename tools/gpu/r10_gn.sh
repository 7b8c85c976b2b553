#!/bin/bash
O=gpurun_out/r10; mkdir -p $O
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "stats or gn or bias" > $O/gn.log 2>&1; echo "gn exit $?" | tee $O/summary.txt; tail -15 $O/gn.log
timeout 1200 python -m pytest tests -q -m gpu > $O/tests.log 2>&1; echo "tests exit $?" | tee -a $O/summary.txt; tail -8 $O/tests.log
timeout 600 python tools/profile_plan.py 32 > $O/plan_B32.csv 2> $O/plan_B32.err; echo "plan exit $?" | tee -a $O/summary.txt; tail -3 $O/plan_B32.err
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu > $O/bench.json 2> $O/bench.err; echo "bench exit $?" | tee -a $O/summary.txt; cat $O/bench.json

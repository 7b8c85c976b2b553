#!/bin/bash
O=gpurun_out/r2_25; mkdir -p $O; rm -f gpurun_out/test_metrics.jsonl
step() { local name=$1 to=$2; shift 2; timeout $to "$@" > $O/$name.log 2>&1; local rc=$?; echo "$name exit $rc" | tee -a $O/summary.txt; tail -6 $O/$name.log; return $rc; }
step kernels 400 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x || exit 0
step smoke 200 python __graft_entry__.py --smoke
step tests 1200 python -m pytest tests -q -m gpu -x --deselect tests/test_kernels_gpu.py
timeout 300 python tools/profile_plan.py 32 > $O/plan.csv 2> $O/plan.err; tail -2 $O/plan.err
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu > $O/bench.json 2> $O/bench.err; echo "bench exit $?" | tee -a $O/summary.txt; cut -c1-300 $O/bench.json; tail -3 $O/bench.err
cat gpurun_out/test_metrics.jsonl

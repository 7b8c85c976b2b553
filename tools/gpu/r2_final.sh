#!/bin/bash
O=gpurun_out/r2_final; mkdir -p $O; rm -f gpurun_out/test_metrics.jsonl
step() { local name=$1 to=$2; shift 2; timeout $to "$@" > $O/$name.log 2>&1; local rc=$?; echo "$name exit $rc" | tee -a $O/summary.txt; tail -4 $O/$name.log; return $rc; }
step tests 1800 python -m pytest tests -q -m gpu -x
step smoke 200 python __graft_entry__.py --smoke
timeout 900 python bench.py --steps 50 --warmup 5 > $O/bench.json 2> $O/bench.err; echo "bench exit $?" | tee -a $O/summary.txt; cut -c1-250 $O/bench.json
cat gpurun_out/test_metrics.jsonl

#!/bin/bash
O=gpurun_out/r2_04; mkdir -p $O; rm -f gpurun_out/test_metrics.jsonl
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "conv or gn or bias" > $O/conv.log 2>&1; rc=$?
echo "conv tests exit $rc" | tee $O/summary.txt; tail -5 $O/conv.log
timeout 300 python tools/bench_conv.py full > $O/bench_full.txt 2>&1; cat $O/bench_full.txt
LM2A_CONV_CG=1 timeout 300 python tools/bench_conv.py cg1 > $O/bench_cg1.txt 2>&1; cat $O/bench_cg1.txt
if [ $rc -ne 0 ]; then exit 0; fi
timeout 600 python tools/profile_plan.py 32 > $O/plan_B32.csv 2> $O/plan_B32.err; echo "plan exit $?" | tee -a $O/summary.txt; tail -3 $O/plan_B32.err
timeout 900 python -m pytest tests -q -m gpu -x --deselect tests/test_fullsize_gpu.py > $O/tests.log 2>&1; echo "tests exit $?" | tee -a $O/summary.txt; tail -5 $O/tests.log
timeout 900 python bench.py --steps 50 --warmup 5 --no-cpu > $O/bench.json 2> $O/bench.err; echo "bench exit $?" | tee -a $O/summary.txt; cut -c1-400 $O/bench.json; tail -3 $O/bench.err

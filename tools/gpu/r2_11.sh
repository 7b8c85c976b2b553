#!/bin/bash
O=gpurun_out/r2_11; mkdir -p $O
timeout 120 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "conv_k3_groupnorm_operand" > $O/conv0.log 2>&1; rc=$?
echo "xf smoke tests exit $rc" | tee $O/summary.txt; tail -3 $O/conv0.log
if [ $rc -ne 0 ]; then exit 0; fi
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "conv or gn or bias" > $O/conv.log 2>&1; rc=$?
echo "conv tests exit $rc" | tee -a $O/summary.txt; tail -3 $O/conv.log
if [ $rc -ne 0 ]; then exit 0; fi
timeout 300 python tools/profile_plan.py 32 > $O/plan_new.csv 2> $O/plan_new.err; tail -2 $O/plan_new.err
timeout 300 python tools/bench_conv.py full > $O/bench_full.txt 2>&1; cat $O/bench_full.txt
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu > $O/bench.json 2> $O/bench.err; echo "bench exit $?" | tee -a $O/summary.txt; cut -c1-300 $O/bench.json; tail -3 $O/bench.err

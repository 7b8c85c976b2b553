#!/bin/bash
# round 2, call 1: the rewritten conv kernel (shared-tap A blocks, operand transform, exact
# GroupNorm sums) — kernel tests first (both descriptor variants), then the suite, smoke, bench.
O=gpurun_out/r2_01; mkdir -p $O; rm -f gpurun_out/test_metrics.jsonl
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $O/gpu.txt 2>&1
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "conv_k3 or conv_k4s2 or conv_k1 or two_segments" > $O/conv_bo1.log 2>&1; rc=$?
echo "conv (base offset field) exit $rc" | tee $O/summary.txt; tail -5 $O/conv_bo1.log
if [ $rc -ne 0 ]; then
  LM2A_DESC_BASE_OFFSET=0 timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "conv_k3 or conv_k4s2 or conv_k1 or two_segments" > $O/conv_bo0.log 2>&1; rc0=$?
  echo "conv (no base offset) exit $rc0" | tee -a $O/summary.txt; tail -5 $O/conv_bo0.log
  if [ $rc0 -ne 0 ]; then exit 0; fi
  export LM2A_DESC_BASE_OFFSET=0
fi
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -m gpu > $O/kernels.log 2>&1; echo "kernels exit $?" | tee -a $O/summary.txt; tail -15 $O/kernels.log
timeout 300 python __graft_entry__.py --smoke > $O/smoke.log 2>&1; echo "smoke exit $?" | tee -a $O/summary.txt; tail -3 $O/smoke.log
timeout 1500 python -m pytest tests -q -m gpu --deselect tests/test_kernels_gpu.py > $O/tests.log 2>&1; echo "tests exit $?" | tee -a $O/summary.txt; tail -25 $O/tests.log
timeout 600 python tools/profile_plan.py 32 > $O/plan_B32.csv 2> $O/plan_B32.err; echo "plan exit $?" | tee -a $O/summary.txt; tail -3 $O/plan_B32.err
timeout 900 python bench.py --steps 50 --warmup 5 > $O/bench.json 2> $O/bench.err; echo "bench exit $?" | tee -a $O/summary.txt; cut -c1-600 $O/bench.json; tail -3 $O/bench.err
cat gpurun_out/test_metrics.jsonl 2>/dev/null

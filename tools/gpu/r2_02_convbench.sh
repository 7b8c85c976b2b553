#!/bin/bash
# conv microbenchmarks: shared-tap A blocks vs one box per tap, cost of the row-shifted views,
# cost of the operand transform; then kernel tests + suite + plan profile + bench
O=gpurun_out/r2_02; mkdir -p $O; rm -f gpurun_out/test_metrics.jsonl
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "conv" > $O/conv.log 2>&1; rc=$?
echo "conv tests exit $rc" | tee $O/summary.txt; tail -5 $O/conv.log
if [ $rc -ne 0 ]; then exit 0; fi
LM2A_CONV_SHARE_TAPS=0 timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "conv" > $O/conv_noshare.log 2>&1; echo "conv tests (per-tap boxes) exit $?" | tee -a $O/summary.txt; tail -3 $O/conv_noshare.log
timeout 300 python tools/bench_conv.py share > $O/bench_share.txt 2>&1; cat $O/bench_share.txt
LM2A_CONV_SHARE_TAPS=0 timeout 300 python tools/bench_conv.py pertap > $O/bench_pertap.txt 2>&1; cat $O/bench_pertap.txt
LM2A_CONV_DBG_NOSHIFT=1 timeout 300 python tools/bench_conv.py noshift > $O/bench_noshift.txt 2>&1; cat $O/bench_noshift.txt
timeout 600 python tools/profile_plan.py 32 > $O/plan_B32.csv 2> $O/plan_B32.err; echo "plan exit $?" | tee -a $O/summary.txt; tail -3 $O/plan_B32.err
LM2A_CONV_SHARE_TAPS=0 timeout 600 python tools/profile_plan.py 32 > $O/plan_B32_pertap.csv 2> $O/plan_B32_pertap.err; tail -3 $O/plan_B32_pertap.err
timeout 900 python -m pytest tests -q -m gpu -x --deselect tests/test_fullsize_gpu.py > $O/tests.log 2>&1; echo "tests exit $?" | tee -a $O/summary.txt; tail -5 $O/tests.log
timeout 900 python bench.py --steps 50 --warmup 5 --no-cpu > $O/bench.json 2> $O/bench.err; echo "bench exit $?" | tee -a $O/summary.txt; cut -c1-400 $O/bench.json; tail -3 $O/bench.err

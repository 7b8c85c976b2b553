#!/bin/bash
# last pass of the round: full GPU suite, smoke, the bench line and the reference arm
O=gpurun_out/r2_final3; mkdir -p $O
timeout 1500 python -m pytest tests -q -m gpu -x > $O/tests.log 2>&1; echo "tests exit $?" | tee $O/summary.txt; tail -3 $O/tests.log
timeout 200 python __graft_entry__.py --smoke > $O/smoke.log 2>&1; echo "smoke exit $?" | tee -a $O/summary.txt; tail -1 $O/smoke.log
timeout 900 python bench.py --steps 50 --warmup 5 > $O/bench_1gpu.json 2> $O/bench_1gpu.err; echo "bench exit $?" | tee -a $O/summary.txt; cut -c1-200 $O/bench_1gpu.json
timeout 300 python tools/profile_plan.py 32 > $O/plan_B32.csv 2> $O/plan_B32.err; tail -2 $O/plan_B32.err
for l in 0 1; do timeout 60 python tools/bench_attn.py $l 32 20; done 2>&1 | tee $O/attn_bench.txt
for l in 2 3; do timeout 60 python tools/bench_attn.py $l 32 20 cond; done 2>&1 | tee -a $O/attn_bench.txt

#!/bin/bash
O=gpurun_out/r2_16; mkdir -p $O; rm -f gpurun_out/test_metrics.jsonl
step() { local name=$1 to=$2; shift 2; timeout $to "$@" > $O/$name.log 2>&1; local rc=$?; echo "$name exit $rc" | tee -a $O/summary.txt; tail -3 $O/$name.log; return $rc; }
step xf_gn 120 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "conv_k3_groupnorm_operand or fused_upsampling" || exit 0
step kernels 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x || exit 0
timeout 300 python tools/profile_plan.py 32 > $O/plan_new.csv 2> $O/plan_new.err; tail -2 $O/plan_new.err
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu > $O/bench.json 2> $O/bench.err; echo "bench exit $?" | tee -a $O/summary.txt; cut -c1-300 $O/bench.json; tail -3 $O/bench.err
# same box, round-1 tree (A/B baseline)
( cd tools/ab/r1_tree && timeout 300 python tools/profile_plan.py 32 > ../../../$O/plan_r1.csv 2> ../../../$O/plan_r1.err; tail -2 ../../../$O/plan_r1.err; timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu > ../../../$O/bench_r1.json 2> ../../../$O/bench_r1.err; cut -c1-300 ../../../$O/bench_r1.json )

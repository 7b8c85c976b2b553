#!/bin/bash
# guarded order: each new feature's own test first (short timeouts), stop at the first failure
O=gpurun_out/r2_15; mkdir -p $O; rm -f gpurun_out/test_metrics.jsonl
step() {  # step <name> <timeout> <cmd...>
  local name=$1 to=$2; shift 2
  timeout $to "$@" > $O/$name.log 2>&1; local rc=$?
  echo "$name exit $rc" | tee -a $O/summary.txt; tail -4 $O/$name.log
  return $rc
}
step xf_gn 120 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "conv_k3_groupnorm_operand" || exit 0
step xf_up 120 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "fused_upsampling" || exit 0
step attn 200 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "cross_attention" || exit 0
step kernels 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x || exit 0
step stepk 200 python -m pytest tests/test_sampler_gpu.py -q -m gpu -x -k "step_kernel or graph_replay" || exit 0
timeout 300 python tools/profile_plan.py 32 > $O/plan_new.csv 2> $O/plan_new.err; tail -2 $O/plan_new.err
timeout 300 python tools/bench_conv.py full > $O/bench_full.txt 2>&1; cat $O/bench_full.txt
step smoke 200 python __graft_entry__.py --smoke
step tests 900 python -m pytest tests -q -m gpu -x --deselect tests/test_kernels_gpu.py
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu > $O/bench.json 2> $O/bench.err; echo "bench exit $?" | tee -a $O/summary.txt; cut -c1-300 $O/bench.json; tail -3 $O/bench.err
cat gpurun_out/test_metrics.jsonl

#!/bin/bash
O=gpurun_out/r2_20; mkdir -p $O; rm -f gpurun_out/test_metrics.jsonl
step() { local name=$1 to=$2; shift 2; timeout $to "$@" > $O/$name.log 2>&1; local rc=$?; echo "$name exit $rc" | tee -a $O/summary.txt; tail -12 $O/$name.log; return $rc; }
step attn_cond 150 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "cross_attention_cond"
step attn 200 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "cross_attention_core" || exit 0
step unet 400 python -m pytest tests/test_unet_gpu.py tests/test_fullsize_gpu.py -q -m gpu -x || exit 0
timeout 300 python tools/profile_plan.py 32 > $O/plan.csv 2> $O/plan.err; tail -2 $O/plan.err; grep cross_attn $O/plan.csv
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu > $O/bench.json 2> $O/bench.err; echo "bench exit $?" | tee -a $O/summary.txt; cut -c1-300 $O/bench.json; tail -3 $O/bench.err

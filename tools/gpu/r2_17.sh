#!/bin/bash
O=gpurun_out/r2_17; mkdir -p $O; rm -f gpurun_out/test_metrics.jsonl
step() { local name=$1 to=$2; shift 2; timeout $to "$@" > $O/$name.log 2>&1; local rc=$?; echo "$name exit $rc" | tee -a $O/summary.txt; tail -3 $O/$name.log; return $rc; }
step unet 400 python -m pytest tests/test_unet_gpu.py tests/test_fullsize_gpu.py -q -m gpu -x || exit 0
for c in 512 256 1024; do
  LM2A_XF_MAX_C=$c LM2A_UP_XF_MAX_C=$c timeout 300 python tools/profile_plan.py 32 > $O/plan_xf$c.csv 2> $O/plan_xf$c.err; echo "xf<=$c"; tail -2 $O/plan_xf$c.err
done
LM2A_XF_MAX_C=512 LM2A_UP_XF_MAX_C=256 timeout 300 python tools/profile_plan.py 32 > $O/plan_xf512_up256.csv 2> $O/plan_xf512_up256.err; tail -2 $O/plan_xf512_up256.err
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu > $O/bench.json 2> $O/bench.err; echo "bench exit $?" | tee -a $O/summary.txt; cut -c1-300 $O/bench.json; tail -3 $O/bench.err
LM2A_XF_MAX_C=256 LM2A_UP_XF_MAX_C=256 timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu > $O/bench_256.json 2> $O/bench_256.err; cut -c1-300 $O/bench_256.json

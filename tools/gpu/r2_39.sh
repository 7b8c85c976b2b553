#!/bin/bash
# gn_apply as one run of slots per CTA with a window of loads in flight
O=gpurun_out/r2_39; mkdir -p $O
step() { local name=$1 to=$2; shift 2; timeout $to "$@" > $O/$name.log 2>&1; local rc=$?; echo "$name exit $rc" | tee -a $O/summary.txt; tail -4 $O/$name.log; return $rc; }
step gn_tests 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "gn or groupnorm or bias_add or stats" || { tail -30 $O/gn_tests.log; exit 0; }
timeout 300 python tools/profile_plan.py 32 > $O/plan_B32.csv 2> $O/plan.err; tail -3 $O/plan.err; grep gn_apply $O/plan_B32.csv | head -30
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu > $O/bench.json 2> $O/bench.err; echo "bench: $(cut -c1-200 $O/bench.json)"
step unet_tests 900 python -m pytest tests/test_unet_gpu.py tests/test_legacy_gpu.py -q -m gpu -x || { tail -30 $O/unet_tests.log; exit 0; }

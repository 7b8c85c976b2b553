#!/bin/bash
O=gpurun_out/r2_30; mkdir -p $O
step() { local name=$1 to=$2; shift 2; timeout $to "$@" > $O/$name.log 2>&1; local rc=$?; echo "$name exit $rc" | tee -a $O/summary.txt; tail -6 $O/$name.log; return $rc; }
LM2A_CONV_SHARE_PLAIN=1 step convs 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "conv_k3 or conv_k4s2 or conv_k1" || exit 0
timeout 200 python tools/bench_conv.py base raw 2>&1 | tee $O/conv_base.txt
LM2A_CONV_SHARE_PLAIN=1 timeout 200 python tools/bench_conv.py share raw 2>&1 | tee $O/conv_share.txt
LM2A_CONV_SHARE_PLAIN=1 timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu > $O/bench_share.json 2> $O/bench_share.err; cut -c1-250 $O/bench_share.json

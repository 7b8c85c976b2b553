#!/bin/bash
O=gpurun_out/r2_28; mkdir -p $O
step() { local name=$1 to=$2; shift 2; timeout $to "$@" > $O/$name.log 2>&1; local rc=$?; echo "$name exit $rc" | tee -a $O/summary.txt; tail -8 $O/$name.log; return $rc; }
step split 200 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "split_k" || exit 0
timeout 200 python tools/bench_conv.py base M2080 2>&1 | tee $O/conv_base.txt
BENCH_SPLIT=1 timeout 200 python tools/bench_conv.py split M2080 2>&1 | tee $O/conv_split.txt
step convs 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "conv" || exit 0
step unet 400 python -m pytest tests/test_unet_gpu.py tests/test_fullsize_gpu.py -q -m gpu -x || exit 0
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu > $O/bench.json 2> $O/bench.err; cut -c1-250 $O/bench.json
LM2A_SPLIT_K=0 timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu > $O/bench_nosplit.json 2> $O/bench_nosplit.err; cut -c1-250 $O/bench_nosplit.json

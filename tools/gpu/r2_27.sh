#!/bin/bash
O=gpurun_out/r2_27; mkdir -p $O
step() { local name=$1 to=$2; shift 2; timeout $to "$@" > $O/$name.log 2>&1; local rc=$?; echo "$name exit $rc" | tee -a $O/summary.txt; tail -6 $O/$name.log; return $rc; }
step attn 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "cross_attention" || exit 0
for l in 0 1; do LM2A_ATTN_RESIDENT=1 timeout 60 python tools/bench_attn.py $l 32 20; timeout 60 python tools/bench_attn.py $l 32 20; done 2>&1 | tee $O/bench.txt
for nz in 2 3 5; do LM2A_ATTN_RESIDENT=1 LM2A_ATTN_NZ=$nz timeout 60 python tools/bench_attn.py 0 32 20; done 2>&1 | tee -a $O/bench.txt

#!/bin/bash
O=gpurun_out/r2_23; mkdir -p $O
step() { local name=$1 to=$2; shift 2; timeout $to "$@" > $O/$name.log 2>&1; local rc=$?; echo "$name exit $rc" | tee -a $O/summary.txt; tail -8 $O/$name.log; return $rc; }
step attn 200 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "cross_attention"
for l in 0 1; do timeout 60 python tools/bench_attn.py $l 32 20; done 2>&1 | tee $O/bench.txt
for l in 2 3; do timeout 60 python tools/bench_attn.py $l 32 20 cond; done 2>&1 | tee -a $O/bench.txt
for nz in 2 3 5; do LM2A_ATTN_NZ=$nz timeout 60 python tools/bench_attn.py 0 32 20; done 2>&1 | tee -a $O/bench.txt
for nz in 2 3; do LM2A_ATTN_NZ=$nz timeout 60 python tools/bench_attn.py 1 32 20; done 2>&1 | tee -a $O/bench.txt
for nz in 1 3 5; do LM2A_ATTN_NZ=$nz timeout 60 python tools/bench_attn.py 2 32 20 cond; done 2>&1 | tee -a $O/bench.txt
step fp32 600 python -m pytest tests/test_unet_gpu.py -q -m gpu -x -k "fp32"
timeout 300 ncu --set full --import-source on --clock-control none -k regex:cross_attn_res --launch-skip 3 --launch-count 1 -o $O/attn_res_l0 -f python tools/bench_attn.py 0 32 2 > $O/ncu0.log 2>&1; tail -2 $O/ncu0.log

#!/bin/bash
O=gpurun_out/r2_21; mkdir -p $O
for l in 0 1; do timeout 60 python tools/bench_attn.py $l 32 20; LM2A_ATTN_RESIDENT=0 timeout 60 python tools/bench_attn.py $l 32 20; done 2>&1 | tee $O/bench.txt
for l in 2 3; do timeout 60 python tools/bench_attn.py $l 32 20 cond; timeout 60 python tools/bench_attn.py $l 32 20; done 2>&1 | tee -a $O/bench.txt
for nz in 1 2 3 5; do LM2A_ATTN_NZ=$nz timeout 60 python tools/bench_attn.py 0 32 20; done 2>&1 | tee -a $O/bench.txt
for nz in 1 2 3; do LM2A_ATTN_NZ=$nz timeout 60 python tools/bench_attn.py 1 32 20; done 2>&1 | tee -a $O/bench.txt
for nz in 1 2 3 4; do LM2A_ATTN_NZ=$nz timeout 60 python tools/bench_attn.py 3 32 20 cond; done 2>&1 | tee -a $O/bench.txt
timeout 300 ncu --set full --import-source on --clock-control none -k regex:cross_attn_res --launch-skip 3 --launch-count 1 -o $O/attn_res_l0 -f python tools/bench_attn.py 0 32 2 > $O/ncu0.log 2>&1; tail -2 $O/ncu0.log
timeout 300 ncu --set full --import-source on --clock-control none -k regex:cross_attn_res --launch-skip 3 --launch-count 1 -o $O/attn_res_l3 -f python tools/bench_attn.py 3 32 2 cond > $O/ncu3.log 2>&1; tail -2 $O/ncu3.log

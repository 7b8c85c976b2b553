#!/bin/bash
O=gpurun_out/r18; mkdir -p $O
for l in 0 1 2 3; do timeout 120 python tools/bench_attn.py $l 32 20 | tee -a $O/attn_bench.txt; done
NCU="ncu --set full --clock-control none --import-source on"
python tools/bench_attn.py 1 32 2 > $O/plain1.log 2>&1 && timeout 600 $NCU -k regex:cross_attn -s 3 -c 1 -f -o $O/attn_tc_l1 python tools/bench_attn.py 1 32 2 > $O/ncu1.log 2>&1
python tools/bench_attn.py 3 32 2 > $O/plain3.log 2>&1 && timeout 600 $NCU -k regex:cross_attn -s 3 -c 1 -f -o $O/attn_tc_l3 python tools/bench_attn.py 3 32 2 > $O/ncu3.log 2>&1
ls -la $O

#!/bin/bash
# where the streaming attention's cycles go: ncu --set full with source-level stall samples
O=gpurun_out/r2_33; mkdir -p $O
NCU="ncu --set full --clock-control none --import-source on"
export LM2A_LIB_PATH=$PWD/tools/probe/poly0/liblm2a_b200.so
timeout 300 $NCU -k regex:cross_attn_tc --launch-skip 3 --launch-count 1 -f -o $O/attn_l0 python tools/bench_attn.py 0 32 3 > $O/ncu0.log 2>&1; tail -1 $O/ncu0.log
timeout 300 $NCU -k regex:cross_attn_tc --launch-skip 3 --launch-count 1 -f -o $O/attn_l1 python tools/bench_attn.py 1 32 3 > $O/ncu1.log 2>&1; tail -1 $O/ncu1.log
ls -la $O

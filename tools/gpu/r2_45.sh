#!/bin/bash
# streaming attention register split: softmax 136 / others 24 (default) vs 128 / 32
O=gpurun_out/r2_45; mkdir -p $O
for i in 1 2; do for v in default tri128; do
  lib=tools/probe/$v/liblm2a_b200.so; [ $v = default ] && lib=lm2a_b200/liblm2a_b200.so
  for lvl in 0 1; do LM2A_LIB_PATH=$PWD/$lib timeout 100 python tools/bench_attn.py $lvl 32 50 2>&1 | tail -1 | sed "s/^/$v /" | tee -a $O/attn.txt; done
done; done

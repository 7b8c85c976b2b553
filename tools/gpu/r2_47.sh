#!/bin/bash
# condition-slab attention register split: softmax 216 / others 72 (default) vs 208 / 88
O=gpurun_out/r2_47; mkdir -p $O
for i in 1 2; do for v in default res208; do
  lib=tools/probe/$v/liblm2a_b200.so; [ $v = default ] && lib=lm2a_b200/liblm2a_b200.so
  for lvl in 2 3; do LM2A_LIB_PATH=$PWD/$lib timeout 100 python tools/bench_attn.py $lvl 32 50 cond 2>&1 | tail -1 | sed "s/^/$v /" | tee -a $O/attn.txt; done
done; done
LM2A_LIB_PATH=$PWD/tools/probe/res208/liblm2a_b200.so timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "attention_cond" 2>&1 | tail -2

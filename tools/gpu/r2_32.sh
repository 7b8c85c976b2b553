#!/bin/bash
# softmax exp2 split over XU / FMA pipes (LM2A_SOFTMAX_POLY = 0..3 pairs of 4) and the
# in-kernel GroupNorm policy for small-M launches (LM2A_XF_MAX_M): tests + A/B
O=gpurun_out/r2_32; mkdir -p $O
step() { local name=$1 to=$2; shift 2; timeout $to "$@" > $O/$name.log 2>&1; local rc=$?; echo "$name exit $rc" | tee -a $O/summary.txt; tail -4 $O/$name.log; return $rc; }
step tests 1500 python -m pytest tests -q -m gpu -x
for k in 0 1 2 3; do
  lib=tools/probe/poly$k/liblm2a_b200.so; [ $k = 1 ] && lib=lm2a_b200/liblm2a_b200.so
  for lvl in 0 1; do LM2A_LIB_PATH=$PWD/$lib timeout 100 python tools/bench_attn.py $lvl 32 50 2>&1 | tail -1 | sed "s/^/poly$k /" | tee -a $O/attn.txt; done
  for lvl in 2 3; do LM2A_LIB_PATH=$PWD/$lib timeout 100 python tools/bench_attn.py $lvl 32 50 cond 2>&1 | tail -1 | sed "s/^/poly$k cond /" | tee -a $O/attn.txt; done
done
for k in 0 2; do
  LM2A_LIB_PATH=$PWD/tools/probe/poly$k/liblm2a_b200.so timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu > $O/bench_poly$k.json 2> $O/bench_poly$k.err; echo "poly$k: $(cut -c1-200 $O/bench_poly$k.json)"
done
for m in 0 2080 4160; do
  LM2A_XF_MAX_M=$m timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu > $O/bench_xfm$m.json 2> $O/bench_xfm$m.err; echo "xfm$m: $(cut -c1-200 $O/bench_xfm$m.json)"
done

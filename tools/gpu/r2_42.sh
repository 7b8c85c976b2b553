#!/bin/bash
# condition-slab attention: 12-warp CTA with setmaxnreg re-balancing (softmax 216 registers)
O=gpurun_out/r2_42; mkdir -p $O
step() { local name=$1 to=$2; shift 2; timeout $to "$@" > $O/$name.log 2>&1; local rc=$?; echo "$name exit $rc" | tee -a $O/summary.txt; tail -4 $O/$name.log; return $rc; }
step attn_tests 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "attention" || { tail -30 $O/attn_tests.log; exit 0; }
for v in new old new old; do
  lib=lm2a_b200/liblm2a_b200.so; [ $v = old ] && lib=tools/probe/res_old/liblm2a_b200.so
  for lvl in 2 3; do LM2A_LIB_PATH=$PWD/$lib timeout 100 python tools/bench_attn.py $lvl 32 50 cond 2>&1 | tail -1 | sed "s/^/$v cond /" | tee -a $O/attn.txt; done
done
step unet_tests 900 python -m pytest tests/test_unet_gpu.py tests/test_sampler_gpu.py -q -m gpu -x || { tail -30 $O/unet_tests.log; exit 0; }
for v in new old; do
  lib=lm2a_b200/liblm2a_b200.so; [ $v = old ] && lib=tools/probe/res_old/liblm2a_b200.so
  LM2A_LIB_PATH=$PWD/$lib timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu > $O/bench_$v.json 2> $O/bench_$v.err; echo "$v: $(cut -c1-200 $O/bench_$v.json)"
done

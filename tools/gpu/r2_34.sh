#!/bin/bash
# d_h = 32 attention with three CTAs per SM (setmaxnreg) x exp2 on the FMA pipe: tests + A/B
O=gpurun_out/r2_34; mkdir -p $O
step() { local name=$1 to=$2; shift 2; timeout $to "$@" > $O/$name.log 2>&1; local rc=$?; echo "$name exit $rc" | tee -a $O/summary.txt; tail -4 $O/$name.log; return $rc; }
step attn_tests 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "attn or attention" || exit 0
for v in tri0poly0 tri1poly0 default tri1poly2; do
  lib=tools/probe/$v/liblm2a_b200.so; [ $v = default ] && lib=lm2a_b200/liblm2a_b200.so
  for lvl in 0 1; do LM2A_LIB_PATH=$PWD/$lib timeout 100 python tools/bench_attn.py $lvl 32 50 2>&1 | tail -1 | sed "s/^/$v /" | tee -a $O/attn.txt; done
done
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu > $O/bench.json 2> $O/bench.err; echo "default: $(cut -c1-200 $O/bench.json)"
step unet_tests 900 python -m pytest tests/test_unet_gpu.py tests/test_sampler_gpu.py -q -m gpu -x

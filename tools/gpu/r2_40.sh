#!/bin/bash
# conv epilogue: next chunk's accumulator + residual fetched ahead (plain launches) - tests + A/B
O=gpurun_out/r2_40; mkdir -p $O
step() { local name=$1 to=$2; shift 2; timeout $to "$@" > $O/$name.log 2>&1; local rc=$?; echo "$name exit $rc" | tee -a $O/summary.txt; tail -4 $O/$name.log; return $rc; }
step conv_tests 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "conv" || { tail -30 $O/conv_tests.log; exit 0; }
step unet_tests 900 python -m pytest tests/test_unet_gpu.py tests/test_sampler_gpu.py tests/test_legacy_gpu.py -q -m gpu -x || { tail -30 $O/unet_tests.log; exit 0; }
timeout 200 python tools/bench_conv.py new raw 2>&1 | tee $O/conv_new.txt
LM2A_LIB_PATH=$PWD/tools/probe/conv_old/liblm2a_b200.so timeout 200 python tools/bench_conv.py old raw 2>&1 | tee $O/conv_old.txt
for v in new old new old; do
  lib=lm2a_b200/liblm2a_b200.so; [ $v = old ] && lib=tools/probe/conv_old/liblm2a_b200.so
  LM2A_LIB_PATH=$PWD/$lib timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu > $O/bench_$v.json 2> $O/bench_$v.err; echo "$v: $(cut -c1-200 $O/bench_$v.json)"
done

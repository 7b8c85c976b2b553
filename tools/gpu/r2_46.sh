#!/bin/bash
# streaming attention d_h = 32: softmax / other registers 128 / 32 (new default) vs 120 / 40; tests + bench
O=gpurun_out/r2_46; mkdir -p $O
step() { local name=$1 to=$2; shift 2; timeout $to "$@" > $O/$name.log 2>&1; local rc=$?; echo "$name exit $rc" | tee -a $O/summary.txt; tail -3 $O/$name.log; return $rc; }
step attn_tests 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "attention" || { tail -30 $O/attn_tests.log; exit 0; }
for i in 1 2; do for v in default tri120; do
  lib=tools/probe/$v/liblm2a_b200.so; [ $v = default ] && lib=lm2a_b200/liblm2a_b200.so
  LM2A_LIB_PATH=$PWD/$lib timeout 100 python tools/bench_attn.py 0 32 50 2>&1 | tail -1 | sed "s/^/$v /" | tee -a $O/attn.txt
done; done
timeout 100 python tools/bench_attn.py 1 32 50 2>&1 | tail -1 | tee -a $O/attn.txt
step unet_tests 900 python -m pytest tests/test_unet_gpu.py tests/test_sampler_gpu.py -q -m gpu -x || { tail -30 $O/unet_tests.log; exit 0; }
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu > $O/bench.json 2> $O/bench.err; echo "bench: $(cut -c1-200 $O/bench.json)"

#!/bin/bash
O=gpurun_out/r2_26; mkdir -p $O
step() { local name=$1 to=$2; shift 2; timeout $to "$@" > $O/$name.log 2>&1; local rc=$?; echo "$name exit $rc" | tee -a $O/summary.txt; tail -6 $O/$name.log; return $rc; }
step attn 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "cross_attention" || exit 0
for l in 0 1 2 3; do timeout 60 python tools/bench_attn.py $l 32 20; done 2>&1 | tee $O/bench.txt
step legacy 300 python -m pytest tests/test_legacy_gpu.py tests/test_unet_gpu.py -q -m gpu -x
for c in 512 256; do
  LM2A_XF_MAX_C=$c LM2A_UP_XF_MAX_C=$c timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu > $O/bench_xf$c.json 2> $O/bench_xf$c.err; cut -c1-250 $O/bench_xf$c.json
done
LM2A_XF_MAX_C=256 LM2A_UP_XF_MAX_C=512 timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu > $O/bench_xf256_up512.json 2> $O/bench.err; cut -c1-250 $O/bench_xf256_up512.json
for m in 0 1; do LM2A_CONV_DBG_STATS=$m timeout 200 python tools/bench_conv.py stats$m raw; done 2>&1 | tee $O/conv_stats.txt

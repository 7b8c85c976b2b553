#!/bin/bash
# condition-slab attention: tail tiles in a CTA of their own (split_tail); tail kernel opt-in
O=gpurun_out/r2_36; mkdir -p $O
step() { local name=$1 to=$2; shift 2; timeout $to "$@" > $O/$name.log 2>&1; local rc=$?; echo "$name exit $rc" | tee -a $O/summary.txt; tail -4 $O/$name.log; return $rc; }
step attn_tests 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "attention" || { tail -30 $O/attn_tests.log; exit 0; }
for sp in 0 1; do for lvl in 2 3; do LM2A_ATTN_SPLIT_TAIL=$sp timeout 100 python tools/bench_attn.py $lvl 32 50 cond 2>&1 | tail -1 | sed "s/^/split$sp cond /" | tee -a $O/attn.txt; done; done
LM2A_ATTN_SPLIT_TAIL=1 timeout 100 python tools/bench_attn.py 2 64 50 cond 2>&1 | tail -1 | sed "s/^/split1 B64 cond /" | tee -a $O/attn.txt
LM2A_ATTN_SPLIT_TAIL=0 timeout 100 python tools/bench_attn.py 2 64 50 cond 2>&1 | tail -1 | sed "s/^/split0 B64 cond /" | tee -a $O/attn.txt
step unet_tests 900 python -m pytest tests/test_unet_gpu.py tests/test_sampler_gpu.py -q -m gpu -x || { tail -30 $O/unet_tests.log; exit 0; }
for v in 1 0 1 0; do
  LM2A_ATTN_SPLIT_TAIL=$v timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu > $O/bench_split$v.json 2> $O/bench_split$v.err; echo "split$v: $(cut -c1-200 $O/bench_split$v.json)"
done

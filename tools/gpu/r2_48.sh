#!/bin/bash
# streaming attention: lean barrier waits in the producer / issuer roles, 16-column rescale path
O=gpurun_out/r2_48; mkdir -p $O
step() { local name=$1 to=$2; shift 2; timeout $to "$@" > $O/$name.log 2>&1; local rc=$?; echo "$name exit $rc" | tee -a $O/summary.txt; tail -3 $O/$name.log; return $rc; }
step attn_tests 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "attention" || { tail -30 $O/attn_tests.log; exit 0; }
for i in 1 2; do for lvl in 0 1; do timeout 100 python tools/bench_attn.py $lvl 32 50 2>&1 | tail -1 | tee -a $O/attn.txt; done; done
step unet_tests 900 python -m pytest tests/test_unet_gpu.py tests/test_sampler_gpu.py tests/test_legacy_gpu.py tests/test_fullsize_gpu.py -q -m gpu -x || { tail -30 $O/unet_tests.log; exit 0; }
timeout 600 python bench.py --steps 50 --warmup 5 > $O/bench_1gpu.json 2> $O/bench.err; echo "bench: $(cut -c1-200 $O/bench_1gpu.json)"

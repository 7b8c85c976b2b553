#!/bin/bash
O=gpurun_out/r17; mkdir -p $O
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "cross_attention" > $O/attn.log 2>&1; echo "attn exit $?" | tee $O/summary.txt; tail -5 $O/attn.log
for l in 0 1 2 3; do timeout 120 python tools/bench_attn.py $l 32 20 | tee -a $O/attn_bench.txt; done

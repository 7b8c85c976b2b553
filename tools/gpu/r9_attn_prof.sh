#!/bin/bash
O=gpurun_out/r9; mkdir -p $O
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "cross_attention" > $O/attn.log 2>&1; echo "attn exit $?" | tee $O/summary.txt; tail -5 $O/attn.log
for l in 0 1 2 3; do python tools/bench_attn.py $l 32 20 | tee -a $O/attn_bench.txt; done
NCU="ncu --set full --clock-control none --import-source on"
python tools/bench_attn.py 0 32 2 > $O/plain0.log 2>&1 && timeout 600 $NCU -k regex:cross_attn -s 3 -c 1 -f -o $O/attn_tc_l0 python tools/bench_attn.py 0 32 2 > $O/ncu0.log 2>&1
python tools/bench_attn.py 2 32 2 > $O/plain2.log 2>&1 && timeout 600 $NCU -k regex:cross_attn -s 3 -c 1 -f -o $O/attn_tc_l2 python tools/bench_attn.py 2 32 2 > $O/ncu2.log 2>&1
ls -la $O

#!/bin/bash
O=gpurun_out/r21; mkdir -p $O
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "cross_attention" > $O/attn.log 2>&1; rc=$?; echo "attn exit $rc" | tee $O/summary.txt; tail -12 $O/attn.log
if [ $rc -ne 0 ]; then exit 0; fi
for l in 0 1 2 3; do timeout 120 python tools/bench_attn.py $l 32 20 | tee -a $O/attn_bench.txt; done
timeout 900 python -m pytest tests -q -m gpu -x > $O/tests.log 2>&1; echo "tests exit $?" | tee -a $O/summary.txt; tail -4 $O/tests.log
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu > $O/bench.json 2> $O/bench.err; echo "bench exit $?" | tee -a $O/summary.txt; cut -c1-250 $O/bench.json

#!/bin/bash
O=gpurun_out/r23; mkdir -p $O
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "fused_groupnorm" > $O/fgn.log 2>&1; rc=$?; echo "fgn exit $rc" | tee $O/summary.txt; tail -25 $O/fgn.log
if [ $rc -ne 0 ]; then exit 0; fi
timeout 900 python -m pytest tests -q -m gpu -x > $O/tests.log 2>&1; echo "tests exit $?" | tee -a $O/summary.txt; tail -4 $O/tests.log
timeout 600 python tools/profile_plan.py 32 > $O/plan.csv 2> $O/plan.err; tail -2 $O/plan.err
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu > $O/bench.json 2> $O/bench.err; echo "bench exit $?" | tee -a $O/summary.txt; cut -c1-250 $O/bench.json; tail -3 $O/bench.err
LM2A_FUSE_GN=0 timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu > $O/bench_nofuse.json 2> $O/bench_nofuse.err; cut -c1-250 $O/bench_nofuse.json

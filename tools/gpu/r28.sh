#!/bin/bash
O=gpurun_out/r28; mkdir -p $O
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "gn" > $O/gn.log 2>&1; echo "gn exit $?" | tee $O/summary.txt; tail -3 $O/gn.log
timeout 600 python tools/profile_plan.py 32 > $O/plan.csv 2> $O/plan.err; tail -2 $O/plan.err; grep gn_apply $O/plan.csv | cut -d, -f1,7 | tr '\n' ' '

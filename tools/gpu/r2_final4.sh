#!/bin/bash
O=gpurun_out/r2_final4; mkdir -p $O
timeout 1500 python -m pytest tests -q -m gpu -x > $O/tests.log 2>&1; echo "tests exit $?" | tee $O/summary.txt; tail -3 $O/tests.log
timeout 200 python __graft_entry__.py --smoke > $O/smoke.log 2>&1; echo "smoke exit $?" | tee -a $O/summary.txt; tail -1 $O/smoke.log

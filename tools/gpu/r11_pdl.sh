#!/bin/bash
O=gpurun_out/r11; mkdir -p $O
timeout 1200 python -m pytest tests -q -m gpu -x > $O/tests.log 2>&1; echo "tests exit $?" | tee $O/summary.txt; tail -8 $O/tests.log
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu > $O/bench_pdl.json 2> $O/bench_pdl.err; echo "bench pdl exit $?" | tee -a $O/summary.txt; cut -c1-200 $O/bench_pdl.json
LM2A_PDL=0 timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu > $O/bench_nopdl.json 2> $O/bench_nopdl.err; echo "bench nopdl exit $?" | tee -a $O/summary.txt; cut -c1-200 $O/bench_nopdl.json
tail -3 $O/bench_pdl.err

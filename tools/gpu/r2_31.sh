#!/bin/bash
O=gpurun_out/r2_31; mkdir -p $O
step() { local name=$1 to=$2; shift 2; timeout $to "$@" > $O/$name.log 2>&1; local rc=$?; echo "$name exit $rc" | tee -a $O/summary.txt; tail -6 $O/$name.log; return $rc; }
step sampler 600 python -m pytest tests/test_sampler_gpu.py -q -m gpu -x || exit 0
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu > $O/bench.json 2> $O/bench.err; cut -c1-250 $O/bench.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:cfg_step --launch-skip 2 --launch-count 3 python bench.py --steps 3 --warmup 3 --no-cpu 2>&1 | grep -E "cfg_step|duration" | head -8

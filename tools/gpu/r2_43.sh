#!/bin/bash
# condition-slab attention: O row into registers, o_free before the stores
O=gpurun_out/r2_43; mkdir -p $O
step() { local name=$1 to=$2; shift 2; timeout $to "$@" > $O/$name.log 2>&1; local rc=$?; echo "$name exit $rc" | tee -a $O/summary.txt; tail -4 $O/$name.log; return $rc; }
step attn_tests 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "attention" || { tail -30 $O/attn_tests.log; exit 0; }
for i in 1 2; do for lvl in 2 3; do timeout 100 python tools/bench_attn.py $lvl 32 50 cond 2>&1 | tail -1 | sed "s/^/earlyfree cond /" | tee -a $O/attn.txt; done; done
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu > $O/bench.json 2> $O/bench.err; echo "bench: $(cut -c1-200 $O/bench.json)"

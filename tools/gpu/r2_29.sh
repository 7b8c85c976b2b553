#!/bin/bash
O=gpurun_out/r2_29; mkdir -p $O
step() { local name=$1 to=$2; shift 2; timeout $to "$@" > $O/$name.log 2>&1; local rc=$?; echo "$name exit $rc" | tee -a $O/summary.txt; tail -6 $O/$name.log; return $rc; }
step convs 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "conv" || exit 0
timeout 200 python tools/bench_conv.py lean raw 2>&1 | tee $O/conv_lean.txt
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu > $O/bench.json 2> $O/bench.err; cut -c1-250 $O/bench.json
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 $NCU -k regex:conv_gemm --launch-skip 5 --launch-count 1 -f -o $O/conv_m8320_n1024_k3584 python tools/bench_conv.py ev "l2 conv2 1024 k3+skip512 raw" > $O/ncu1.log 2>&1; tail -1 $O/ncu1.log
timeout 300 $NCU -k regex:conv_gemm --launch-skip 5 --launch-count 1 -f -o $O/conv_xf_m16640_n512_k768 python tools/bench_conv.py ev "l0 q-proj 256->512 k3" > $O/ncu2.log 2>&1; tail -1 $O/ncu2.log

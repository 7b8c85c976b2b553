#!/bin/bash
O=gpurun_out/r2_05; mkdir -p $O
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "conv or gn or bias" > $O/conv.log 2>&1; rc=$?
echo "conv tests exit $rc" | tee $O/summary.txt; tail -3 $O/conv.log
timeout 300 python tools/bench_conv.py full "l3 conv1" > $O/bench_full.txt 2>&1; cat $O/bench_full.txt
LM2A_CONV_DBG_NOXFORM=1 timeout 300 python tools/bench_conv.py relay_only "l3 conv1" > $O/bench_relay.txt 2>&1; cat $O/bench_relay.txt
BENCH_GN_SILU=0 timeout 300 python tools/bench_conv.py affine_only "l3 conv1" > $O/bench_affine.txt 2>&1; cat $O/bench_affine.txt
python tools/bench_conv.py ncu "l3 conv1 1024 k3" > $O/plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_gemm -s 12 -c 1 -f -o $O/conv_xf_l3 python tools/bench_conv.py ncu "l3 conv1 1024 k3" > $O/ncu.log 2>&1; echo "ncu exit $?"; tail -2 $O/ncu.log

#!/bin/bash
# where does the operand transform's time go: pass-through relay only, affine without SiLU,
# full transform; ncu source-level capture of one transform-bound launch
O=gpurun_out/r2_03; mkdir -p $O
timeout 300 python tools/bench_conv.py full > $O/bench_full.txt 2>&1; cat $O/bench_full.txt
LM2A_CONV_DBG_NOXFORM=1 timeout 300 python tools/bench_conv.py relay_only > $O/bench_relay.txt 2>&1; grep -v raw $O/bench_relay.txt
BENCH_GN_SILU=0 timeout 300 python tools/bench_conv.py affine_only > $O/bench_affine.txt 2>&1; grep -v raw $O/bench_affine.txt
python tools/bench_conv.py ncu "l3 conv1 1024 k3" > $O/plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_gemm -s 12 -c 1 -f -o $O/conv_xf_l3 python tools/bench_conv.py ncu "l3 conv1 1024 k3" > $O/ncu.log 2>&1; echo "ncu exit $?"; tail -3 $O/ncu.log
ls -la $O

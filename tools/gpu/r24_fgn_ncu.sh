#!/bin/bash
O=gpurun_out/r24; mkdir -p $O
python tools/run_step.py 32 2 > $O/plain.log 2>&1 || exit 1
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:conv_gemm -s 19 -c 1 -f -o $O/conv1_fused_l0 python tools/run_step.py 32 2 > $O/ncu1.log 2>&1
ls -la $O

#!/bin/bash
O=gpurun_out/r14; mkdir -p $O
timeout 600 python tools/bench_conv.py > $O/conv_bench.txt 2>&1; cat $O/conv_bench.txt

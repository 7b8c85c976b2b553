#!/bin/bash
O=gpurun_out/r16; mkdir -p $O
for s in 1 2 4; do timeout 300 python tools/exp_two_streams.py 32 $s 2>&1 | tail -2 | tee -a $O/streams.txt; done
for s in 2; do timeout 300 python tools/exp_two_streams.py 64 $s 2>&1 | tail -2 | tee -a $O/streams.txt; done
for s in 1; do timeout 300 python tools/exp_two_streams.py 64 $s 2>&1 | tail -2 | tee -a $O/streams.txt; done

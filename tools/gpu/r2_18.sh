#!/bin/bash
O=gpurun_out/r2_18; mkdir -p $O
timeout 200 python tools/bench_conv.py base raw > $O/base.txt 2>&1; cat $O/base.txt
BENCH_NO_STATS=1 timeout 200 python tools/bench_conv.py nostats raw > $O/nostats.txt 2>&1; cat $O/nostats.txt
LM2A_CONV_TAP_OUTER=1 timeout 200 python tools/bench_conv.py tapouter raw > $O/tapouter.txt 2>&1; cat $O/tapouter.txt
LM2A_CONV_TAP_OUTER=1 BENCH_NO_STATS=1 timeout 200 python tools/bench_conv.py tapouter_nostats raw > $O/tapouter_nostats.txt 2>&1; cat $O/tapouter_nostats.txt
LM2A_CONV_TAP_OUTER=1 timeout 300 python tools/profile_plan.py 32 > $O/plan_tapouter.csv 2> $O/plan_tapouter.err; tail -2 $O/plan_tapouter.err

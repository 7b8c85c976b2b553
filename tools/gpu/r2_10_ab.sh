#!/bin/bash
O=gpurun_out/r2_10; mkdir -p $O
timeout 600 python tools/profile_plan.py 32 > $O/plan_new.csv 2> $O/plan_new.err; tail -2 $O/plan_new.err
LM2A_CONV_SHARE_TAPS=0 LM2A_LIB_PATH=$PWD/tools/ab/liblm2a_r2_02.so timeout 600 python tools/profile_plan.py 32 > $O/plan_old.csv 2> $O/plan_old.err; tail -2 $O/plan_old.err
timeout 600 python tools/profile_plan.py 32 > $O/plan_new2.csv 2> $O/plan_new2.err; tail -2 $O/plan_new2.err

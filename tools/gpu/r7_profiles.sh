#!/bin/bash
# full GPU test pass, bench line, ncu launch list of the bench command, ncu --set full of the
# top kernels (one launch each). Outputs under gpurun_out/r7/.
O=gpurun_out/r7; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $O/gpu.txt 2>&1
timeout 1500 python -m pytest tests -q -m gpu -x > $O/tests.log 2>&1; echo "tests exit $?" | tee $O/summary.txt; tail -3 $O/tests.log
timeout 600 python bench.py --steps 50 --warmup 5 > $O/bench.json 2> $O/bench.err; echo "bench exit $?" | tee -a $O/summary.txt
timeout 300 python bench.py --impl reference --steps 10 --warmup 2 > $O/bench_ref.json 2> $O/bench_ref.err; echo "bench ref exit $?" | tee -a $O/summary.txt
timeout 600 python tools/profile_plan.py 32 > $O/plan_B32.csv 2> $O/plan_B32.err; echo "plan exit $?" | tee -a $O/summary.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu > $O/ncu_bench.log 2>&1; echo "ncu launches exit $?" | tee -a $O/summary.txt
python tools/run_step.py 32 2 > $O/plain.log 2>&1 || exit 1
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:conv_gemm -s 34 -c 1 -f -o $O/conv_m8320_n1024_k3584 python tools/run_step.py 32 2 > $O/ncu1.log 2>&1
timeout 600 $NCU -k regex:conv_gemm -s 40 -c 1 -f -o $O/conv_m2080_n1024_k3072 python tools/run_step.py 32 2 > $O/ncu2.log 2>&1
timeout 600 $NCU -k regex:gn_silu -s 0 -c 1 -f -o $O/gn_l0 python tools/run_step.py 32 2 > $O/ncu3.log 2>&1
timeout 600 $NCU -k regex:cross_attn -s 0 -c 1 -f -o $O/attn_l0 python tools/run_step.py 32 2 > $O/ncu4.log 2>&1
timeout 600 $NCU -k regex:cfg_posterior -s 0 -c 1 -f -o $O/cfg_posterior python tools/run_step.py 32 2 > $O/ncu5.log 2>&1
ls -la $O
cat $O/bench.json

#!/bin/bash
# round-2 final evidence pass (one B200) after the attention changes (three CTAs per SM at
# d_h = 32 / 64): full GPU test suite, smoke, bench line + reference arm, launch list of the
# bench command, DRAM traffic of one eager step, per-launch plan times, ncu --set full of the
# changed attention kernels, the other BASELINE configs. Outputs under gpurun_out/r2_ev2/.
O=gpurun_out/r2_ev2; mkdir -p $O; rm -f gpurun_out/test_metrics.jsonl
K="regex:conv_gemm|cross_attn|gn_apply|gn_silu|film_kernel|time_mlp|ingest|upsample2x|cfg_posterior|cfg_step|philox|bias_add|transpose_kv"
timeout 1800 python -m pytest tests -q -m gpu -x > $O/tests.log 2>&1; echo "tests exit $?" | tee $O/summary.txt; tail -3 $O/tests.log
timeout 200 python __graft_entry__.py --smoke > $O/smoke.log 2>&1; echo "smoke exit $?" | tee -a $O/summary.txt; tail -1 $O/smoke.log
timeout 900 python bench.py --steps 50 --warmup 5 > $O/bench_1gpu.json 2> $O/bench_1gpu.err; echo "bench exit $?" | tee -a $O/summary.txt; cut -c1-200 $O/bench_1gpu.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref exit $?" | tee -a $O/summary.txt; cut -c1-300 $O/bench_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -s 200 -c 400 --csv --log-file $O/launches_bench.csv python bench.py --steps 20 --warmup 3 --no-cpu > $O/ncu_bench.log 2>&1; echo "launch list exit $?" | tee -a $O/summary.txt
python tools/run_step.py 32 2 > $O/plain.log 2>&1 || exit 1
cat $O/plain.log
SKIP=$(sed -n 's/.*skip \([0-9]*\) per_step.*/\1/p' $O/plain.log); PER=$(sed -n 's/.*per_step \([0-9]*\).*/\1/p' $O/plain.log)
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --cache-control all --clock-control none -k "$K" -s $SKIP -c $PER --csv --log-file $O/traffic_step.csv python tools/run_step.py 32 2 > $O/ncu_traffic.log 2>&1; echo "traffic exit $? (skip $SKIP count $PER)" | tee -a $O/summary.txt
timeout 300 python tools/profile_plan.py 32 > $O/plan_B32.csv 2> $O/plan_B32.err; tail -2 $O/plan_B32.err
timeout 300 python tools/profile_plan.py 32 516 tiled > $O/plan_B32_tiled.csv 2> $O/plan_B32_tiled.err; tail -2 $O/plan_B32_tiled.err
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 $NCU -k regex:cross_attn_tc --launch-skip 3 --launch-count 1 -f -o $O/attn_tc_l0_dh32 python tools/bench_attn.py 0 32 2 > $O/ncu5.log 2>&1
timeout 300 $NCU -k regex:cross_attn_tc --launch-skip 3 --launch-count 1 -f -o $O/attn_tc_l1_dh64 python tools/bench_attn.py 1 32 2 > $O/ncu6.log 2>&1
timeout 600 python tools/bench_configs.py 30 > $O/configs.jsonl 2> $O/configs.err; cat $O/configs.jsonl | cut -c1-200
for l in 0 1 2 3; do timeout 60 python tools/bench_attn.py $l 32 20; done 2>&1 | tee $O/attn_bench.txt
for l in 2 3; do timeout 60 python tools/bench_attn.py $l 32 20 cond; done 2>&1 | tee -a $O/attn_bench.txt
cat gpurun_out/test_metrics.jsonl 2>/dev/null | tail -20
ls -la $O

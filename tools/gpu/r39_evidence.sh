#!/bin/bash
# evidence pass: memcheck of smoke, launch list of the bench command, DRAM traffic of every launch
# of one eager step, ncu --set full of the top kernels + the loop-side kernels. -> gpurun_out/r39/
O=gpurun_out/r39; mkdir -p $O
K="regex:conv_gemm|cross_attn|gn_apply|gn_silu|film_kernel|time_mlp|ingest|upsample2x|cfg_posterior|cfg_ddim|bias_add|transpose_kv|resample_seq|mel_metrics"
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python __graft_entry__.py --smoke > $O/memcheck_smoke.log 2>&1; echo "memcheck exit $?" | tee $O/summary.txt
tail -3 $O/memcheck_smoke.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu > $O/bench_plain.json 2> $O/bench_plain.err && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -s 150 -c 400 --csv --log-file $O/launches_bench.csv python bench.py --steps 20 --warmup 3 --no-cpu > $O/ncu_bench.log 2>&1; echo "launch list exit $?" | tee -a $O/summary.txt
python tools/run_step.py 32 2 > $O/plain.log 2>&1 || { cat $O/plain.log; exit 1; }
cat $O/plain.log
SKIP=$(awk '{for(i=1;i<=NF;i++) if($i=="skip") print $(i+1)}' $O/plain.log)
PER=$(awk '{for(i=1;i<=NF;i++) if($i=="per_step") print $(i+1)}' $O/plain.log)
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "$K" -s $SKIP -c $PER --csv --log-file $O/traffic_step.csv python tools/run_step.py 32 2 > $O/ncu_traffic.log 2>&1; echo "traffic exit $? (skip $SKIP, per step $PER)" | tee -a $O/summary.txt
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:conv_gemm -s 32 -c 1 -f -o $O/conv_m8320_n1024_k3584 python tools/run_step.py 32 2 > $O/ncu1.log 2>&1
timeout 600 $NCU -k regex:conv_gemm -s 20 -c 1 -f -o $O/conv_m33280_n256_k768 python tools/run_step.py 32 2 > $O/ncu2.log 2>&1
timeout 600 $NCU -k regex:cross_attn -s 0 -c 1 -f -o $O/attn_l0_dh32 python tools/run_step.py 32 2 > $O/ncu3.log 2>&1
timeout 600 $NCU -k regex:cross_attn -s 2 -c 1 -f -o $O/attn_l2_dh128 python tools/run_step.py 32 2 > $O/ncu4.log 2>&1
timeout 600 $NCU -k regex:gn_apply -s 0 -c 1 -f -o $O/gn_apply_l0 python tools/run_step.py 32 2 > $O/ncu5.log 2>&1
timeout 600 $NCU -k regex:cfg_posterior -s 0 -c 1 -f -o $O/cfg_posterior python tools/run_step.py 32 2 > $O/ncu6.log 2>&1
python tools/run_aux.py > $O/aux_plain.txt 2>&1; cat $O/aux_plain.txt
timeout 300 $NCU -k regex:resample_seq -s 3 -c 1 -f -o $O/resample_motion python tools/run_aux.py > $O/ncu7.log 2>&1
timeout 300 $NCU -k regex:resample_seq -s 26 -c 1 -f -o $O/resample_lyrics python tools/run_aux.py > $O/ncu8.log 2>&1
timeout 300 $NCU -k regex:cfg_ddim -s 3 -c 1 -f -o $O/cfg_ddim python tools/run_aux.py > $O/ncu9.log 2>&1
timeout 300 $NCU -k regex:mel_metrics -s 3 -c 1 -f -o $O/mel_metrics python tools/run_aux.py > $O/ncu10.log 2>&1
ls -la $O | head -40

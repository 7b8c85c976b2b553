#!/bin/bash
# evidence pass: launch list of the bench command (our kernels only), DRAM traffic of every launch of
# one eager step, ncu --set full of the top kernels. Outputs under gpurun_out/r20/.
O=gpurun_out/r20; mkdir -p $O
K="regex:conv_gemm|cross_attn|gn_apply|gn_silu|film_kernel|time_mlp|ingest|upsample2x|cfg_posterior|bias_add|transpose_kv"
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu > $O/bench_plain.json 2> $O/bench_plain.err && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -s 150 -c 400 --csv --log-file $O/launches_bench.csv python bench.py --steps 20 --warmup 3 --no-cpu > $O/ncu_bench.log 2>&1; echo "launch list exit $?" | tee $O/summary.txt
python tools/run_step.py 32 2 > $O/plain.log 2>&1 || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "$K" -s 150 -c 113 --csv --log-file $O/traffic_step.csv python tools/run_step.py 32 2 > $O/ncu_traffic.log 2>&1; echo "traffic exit $?" | tee -a $O/summary.txt
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:conv_gemm -s 34 -c 1 -f -o $O/conv_m8320_n1024_k3584 python tools/run_step.py 32 2 > $O/ncu1.log 2>&1
timeout 600 $NCU -k regex:conv_gemm -s 21 -c 1 -f -o $O/conv_m33280_n256_k768 python tools/run_step.py 32 2 > $O/ncu2.log 2>&1
timeout 600 $NCU -k regex:gn_apply -s 0 -c 1 -f -o $O/gn_apply_l0 python tools/run_step.py 32 2 > $O/ncu3.log 2>&1
timeout 600 $NCU -k regex:cfg_posterior -s 0 -c 1 -f -o $O/cfg_posterior python tools/run_step.py 32 2 > $O/ncu5.log 2>&1
ls -la $O

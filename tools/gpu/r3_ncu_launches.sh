#!/bin/bash
mkdir -p gpurun_out
python tools/run_step.py 32 3 > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:conv_gemm|cross_attn|gn_silu|film_kernel|time_mlp|ingest|upsample2x|cfg_posterior" -c 400 --csv --log-file gpurun_out/launches_r1.csv python tools/run_step.py 32 3 > gpurun_out/ncu.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/plain.log gpurun_out/ncu.log; wc -l gpurun_out/launches_r1.csv

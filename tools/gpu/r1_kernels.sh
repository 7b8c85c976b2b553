#!/bin/bash
# first GPU contact: per-kernel parity in separate processes so a trap in one group
# cannot poison the CUDA context of the others
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
for grp in "gn_silu or upsample or ingest or time_mlp or cfg_posterior" "cross_attention" "conv" ; do
  name=$(echo "$grp" | tr ' ' '_' | cut -c1-24)
  timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "$grp" > "gpurun_out/k_${name}.log" 2>&1
  echo "group [$grp] exit $?" | tee -a gpurun_out/summary.txt
  tail -5 "gpurun_out/k_${name}.log"
done
timeout 900 python -m pytest tests/test_unet_gpu.py -q -m gpu > gpurun_out/unet.log 2>&1
echo "unet exit $?" | tee -a gpurun_out/summary.txt
tail -15 gpurun_out/unet.log

#!/bin/bash
# usage: tools/gpu/retry.sh <tag> <timeout_s> <script>   - retries while the pool answers "transient"
tag=$1; to=$2; script=$3
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout $to -- "bash $script" > gpurun_out/call_$tag.log 2>&1
  if ! grep -q "status=transient" gpurun_out/call_$tag.log; then break; fi
  sleep 90
done
grep -v "^\[gpurun\] merged" gpurun_out/call_$tag.log | cut -c1-220

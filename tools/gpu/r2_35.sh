#!/bin/bash
# leftover query rows on the CUDA cores (parallel branch): tests + A/B against LM2A_ATTN_TAIL=0
O=gpurun_out/r2_35; mkdir -p $O
step() { local name=$1 to=$2; shift 2; timeout $to "$@" > $O/$name.log 2>&1; local rc=$?; echo "$name exit $rc" | tee -a $O/summary.txt; tail -4 $O/$name.log; return $rc; }
step tail_tests 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "tail_rows" || { tail -30 $O/tail_tests.log; exit 0; }
step smoke 200 python __graft_entry__.py --smoke
step unet_tests 900 python -m pytest tests/test_unet_gpu.py tests/test_sampler_gpu.py -q -m gpu -x || { tail -30 $O/unet_tests.log; exit 0; }
for v in 1 0 1 0; do
  LM2A_ATTN_TAIL=$v timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu > $O/bench_tail$v.json 2> $O/bench_tail$v.err; echo "tail$v: $(cut -c1-200 $O/bench_tail$v.json)"
done
python - <<'PY'
import json
for v in (1, 0):
    d = json.load(open(f"gpurun_out/r2_35/bench_tail{v}.json"))
    print(v, d["ms_per_step"], d["kernel_ms"], d.get("tiled_lyrics", {}).get("ms_per_step"), d.get("config3_B64", {}).get("ms_per_step"))
PY
LM2A_ATTN_TAIL=1 timeout 300 python tools/profile_plan.py 32 > $O/plan_B32.csv 2> $O/plan.err; tail -3 $O/plan.err

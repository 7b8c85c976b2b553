#!/bin/bash
O=gpurun_out/r2_41; mkdir -p $O
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu > $O/bench.json 2> $O/bench.err; echo "bench: $(cut -c1-200 $O/bench.json)"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_41/bench.json"))
print(d["ms_per_step"], d["roofline"]["frac"], d["kernel_ms"], sum(d["kernel_ms"].values()))
PY
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 $NCU -k regex:cross_attn_res --launch-skip 3 --launch-count 1 -f -o $O/attn_cond_l2 python tools/bench_attn.py 2 32 3 cond > $O/ncu0.log 2>&1; tail -1 $O/ncu0.log

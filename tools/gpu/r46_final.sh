#!/bin/bash
# final single-GPU pass of the round: full GPU test suite, smoke, bench (both arms)
O=gpurun_out/r46; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5 | tee $O/tests.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2 | tee $O/smoke.log
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench exit $?" | tee $O/summary.txt
timeout 600 python bench.py --impl reference --steps 10 --warmup 2 > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref exit $?" | tee -a $O/summary.txt
python - <<'PY'
import json
d = json.load(open("gpurun_out/r46/bench.json"))
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches", "clocks")})
print("e2e", d["e2e"]["value"], "roofline", d["roofline"]["frac"], d["roofline"]["traffic"])
print("tiled", d["tiled_lyrics"]["ms_per_step"], "ddim", d["ddim"]["clips_per_s"])
print("cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"], "eager", d["torch_eager_gpu_baseline"].get("bf16_autocast_clips_per_s"))
r = json.load(open("gpurun_out/r46/bench_ref.json"))
print("ref", r["value"], r["ms_per_step"])
PY
